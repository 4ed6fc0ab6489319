"""Known-answer vectors.

Sources: (1) the comment examples the reference itself contains — splitter ordering keys
(reference src/utils.rs:50-57) and the IGV two-alignment large insertion (src/main.rs:357-365);
(2) SURVEY.md Appendix B, derived by hand from the cited lines (the reference ships no tests
and cannot be run here).  Each entry: (id, params-overrides, record dict, expected lines).
"""
D = dict  # defaults: Q1 F1796 i50 m5 cmin1000 p0.0 k4

FFM = [("20S30M100S", 20), ("50S30M50S", 50), ("90S30M30S", 90), ("100H30M", 0), ("5S10=2X30M", 17),
       ("5S10=2X30=", 47), ("10I5M", 10), ("3D4S5M", 4)]

SA11 = "chr1,101,+,500S10M2490S,1,0;chr3,201,-,400S10M2590S,1,0;chr1,301,+,600S10M2390S,1,0;"

KA = [
    ("KA1", D(), D(contig="chr20", pos=1000, flag=0, mapq=60, cigar="100M60D200M"),
     ["20\t1000\t1100\t1\t20\t1160\t1360\t1\t1"]),
    ("KA2", D(), D(contig="chr20", pos=1000, flag=16, mapq=60, cigar="10S100M75I200M5S"),
     ["20\t1000\t1100\t-1\t20\t1100\t1175\t-1\t1"]),
    ("KA3", D(), D(contig="chr20", pos=1000, flag=0, mapq=60, cigar="100M50D4M70D300M"),
     ["20\t1000\t1100\t1\t20\t1224\t1524\t1\t1"]),
    ("KA4", D(), D(contig="chr20", pos=1000, flag=0, mapq=60, cigar="100M50D5M70D300M"),
     ["20\t1000\t1100\t1\t20\t1150\t1525\t1\t1", "20\t1000\t1155\t1\t20\t1225\t1525\t1\t1"]),
    ("KA5", D(), D(contig="chr20", pos=1000, flag=0, mapq=60, cigar="100M50D2M50D2M50D300M"),
     ["20\t1000\t1100\t1\t20\t1150\t1554\t1\t1", "20\t1000\t1152\t1\t20\t1202\t1554\t1\t1",
      "20\t1000\t1204\t1\t20\t1254\t1554\t1\t1"]),
    ("KA6", D(), D(contig="20", pos=500, flag=0, mapq=60, cigar="50=10X40=60D10N100="),
     ["20\t500\t590\t1\t20\t650\t760\t1\t1"]),
    ("KA7", D(), D(contig="chr1", pos=0, flag=0, mapq=60, cigar="100M50D2M50I100M"),
     ["1\t0\t100\t1\t1\t150\t252\t1\t1", "1\t0\t152\t1\t1\t152\t202\t1\t1"]),
    ("KA8", D(), D(contig="chr20", pos=10000, flag=0, mapq=60, cigar="5000M3000S", sa="chr20,20001,+,5000S3000M,60,12;"),
     ["20\t10000\t15000\t1\t20\t20000\t23000\t1\t1"]),
    ("KA9", D(), D(contig="chr2", pos=10000, flag=0, mapq=60, cigar="5000M3000S", sa="chr10,501,-,5000S3000M,60,3;"),
     ["2\t10000\t15000\t1\t2\t15000\t15000\t1\t1", "10\t500\t3500\t-1\t2\t10000\t15000\t1\t1"]),
    ("KA10", D(), D(contig="chr20", pos=1000, flag=0, mapq=60, cigar="100M60D200M3000S",
                    sa=SA11 + "chr4,401,+,700S10M2290S,1,0;"), []),
    ("KA11", D(), D(contig="chr20", pos=1000, flag=0, mapq=60, cigar="100M60D200M3000S", sa=SA11),
     ["20\t1000\t1360\t1\t3\t200\t210\t-1\t3", "1\t100\t110\t1\t3\t200\t210\t-1\t3", "1\t100\t110\t1\t1\t300\t310\t1\t3",
      "20\t1000\t1100\t1\t20\t1160\t1360\t1\t1"]),
    ("KA12", D(), D(contig="chr2", pos=9832510, flag=0, mapq=60, cigar="32S45339M8269S",
                    sa="chr2,9877576,+,51421S1538M14S,60,191;"),
     ["2\t9832510\t9877575\t1\t2\t9877575\t9877575\t1\t1", "2\t9832510\t9877849\t1\t2\t9877849\t9877849\t1\t1",
      "2\t9832510\t9877849\t1\t2\t9877575\t9879113\t1\t1"]),
    ("KA12p", D(max_pct_overlap=0.8), D(contig="chr2", pos=9832510, flag=0, mapq=60, cigar="32S45339M8269S",
                                        sa="chr2,9877576,+,51421S1538M14S,60,191;"),
     ["2\t9832510\t9877849\t1\t2\t9877575\t9879113\t1\t1"]),
    ("KA13", D(), D(contig="chr2", pos=1000, flag=0, mapq=60, cigar="2000S5000M", sa="chr5,101,+,2000M5000S,60,1;"),
     ["5\t100\t2100\t1\t5\t2100\t2100\t1\t1", "2\t1000\t6000\t1\t5\t100\t2100\t1\t1"]),
    ("KA14", D(), D(contig="chr20", pos=20000, flag=2048, mapq=60, cigar="5000H3000M", sa="chr20,10001,+,5000M3000S,60,7;"),
     ["20\t10000\t15000\t1\t20\t20000\t23000\t1\t1"]),
    ("KA15", D(split_only=True), D(contig="chr20", pos=10000, flag=0, mapq=60, cigar="5000M60D10M3000S",
                                   sa="chr20,20001,+,5000S3000M,60,12;"),
     ["20\t10000\t15070\t1\t20\t20000\t23000\t1\t1"]),
    ("KA16", D(), D(contig="chr20", pos=1000, flag=256, mapq=60, cigar="100M60D200M"), []),
    ("KA17", D(), D(contig="chr20", pos=1000, flag=0, mapq=0, cigar="100M60D200M"), []),
    ("KA18", D(), D(contig="chr1", pos=100, flag=0, mapq=60, cigar="50M50S", sa="chr1,0,+,50S50M,60,0;"),
     ["1\t-1\t49\t1\t1\t100\t150\t1\t1"]),
]

REF_NAMES = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY", "chrM", "20"]
