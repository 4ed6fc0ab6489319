#!/bin/sh
# Parity-pinning kit, step 1: build the UNMODIFIED reference (zhengxinchang/excord-lr, Rust) into oracle/_ref/excord-lr.
#
# This image has no Rust toolchain (no cargo/rustc, no network for crates), so here the script stops with exit 3 and
# parity stays "unpinned" (DESIGN.md 6).  On a machine that has cargo + the crates (clap, rust-htslib, bio-types; htslib's
# C build needs cc, make, zlib, bzip2, lzma headers), it:
#   - copies the reference tree to a scratch directory (the tree may be read-only and has no Cargo.lock),
#   - runs `cargo build --release` there,
#   - puts the binary at oracle/_ref/excord-lr and the Cargo.lock cargo resolved next to it (the dependency versions
#     the pinned outputs were produced with: the reference pins nothing, Cargo.toml says "*").
# Outputs go ONLY under oracle/_ref/ (git-ignored; it travels to the GPU box).  No reference source is copied into the repo.
# Step 2 is tests/test_ref_binary.py: it finds oracle/_ref/excord-lr and diffs it against the oracle and the CLI.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${EXLR_REFERENCE:-/root/reference}
OUT=$HERE/_ref
if ! command -v cargo >/dev/null 2>&1; then
    echo "build_ref.sh: cargo not found -- the Rust reference cannot be built here; parity stays unpinned" >&2
    exit 3
fi
[ -f "$REF/Cargo.toml" ] || { echo "build_ref.sh: no reference tree at $REF (set EXLR_REFERENCE)" >&2; exit 2; }
mkdir -p "$OUT"
SCRATCH=$(mktemp -d "${TMPDIR:-/tmp}/exlr_ref.XXXXXX")
trap 'rm -rf "$SCRATCH"' EXIT
cp -r "$REF"/Cargo.toml "$REF"/src "$SCRATCH"/
[ -f "$REF/Cargo.lock" ] && cp "$REF/Cargo.lock" "$SCRATCH"/
( cd "$SCRATCH" && cargo build --release ${EXLR_CARGO_FLAGS:-} )
cp "$SCRATCH/target/release/excord-lr" "$OUT/excord-lr"
cp "$SCRATCH/Cargo.lock" "$OUT/Cargo.lock"
( cd "$SCRATCH" && cargo --version && rustc --version ) > "$OUT/toolchain.txt" 2>&1 || true
echo "built $OUT/excord-lr ; dependency versions in $OUT/Cargo.lock"
