#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the reference's own CPU path (qr.c: mmqr, explicitQR, dgemm, identity,
# getPanelDims) from where it lies under /root/reference into shared objects in
# oracle/_ref/ (git-ignored; travels to the GPU box).  No reference source is
# copied into the repo: PR/PC are unguarded #defines (qr.c:12-13), so variants
# other than the file's own 4/2 are produced by a sed on the way into gcc's stdin.
#   libref_qr_4_2.so   unmodified defaults (qr.c:12-13)
#   libref_qr_64_4.so  the GPU file's geometry (qr.cu:21-23)
#   libref_qr_64_8.so  the legal geometry for 512x512 with PR=64
set -euo pipefail
REF=${REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/qr.c" ]; then
  echo "build_ref: $REF/qr.c not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
CFLAGS="-O2 -std=c99 -ffp-contract=off -fPIC -shared -w -include $HERE/ref_shim.h -Dmain=ref_main"
for geo in "4 2" "64 4" "64 8"; do
  set -- $geo
  sed -e "s/^#define PR 4\$/#define PR $1/" -e "s/^#define PC 2\$/#define PC $2/" "$REF/qr.c" \
    | gcc $CFLAGS -x c - -o "$OUT/libref_qr_$1_$2.so" -lm
done
echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
