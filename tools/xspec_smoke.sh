#!/bin/bash
# The reference CI's XSPEC smoke test (/root/reference/.github/workflows/test.yml:78-105): load the package, set up
# kconv5*gaussian with the CI's parameter values, exit.  Needs HEASoft's `xspec` on PATH, the package built by
# `make -C xspec package`, a CUDA device, and $KLINE_PROF_DATA_DIR pointing at kerr-transfer-functions.fits (or .klpt).
# usage: tools/xspec_smoke.sh [package_dir]      exit 0 = passed, 77 = skipped (no xspec here)
set -u
PKG=${1:-$(dirname "$0")/../xspec/build}
if ! command -v xspec > /dev/null 2>&1; then
    echo "xspec_smoke: SKIPPED -- no xspec on PATH (HEASoft is not installed in this environment)"
    exit 77
fi
if [ ! -d "$PKG" ]; then
    echo "xspec_smoke: package directory $PKG missing: run make -C xspec package" >&2
    exit 1
fi
XCM=$(mktemp /tmp/klp_smoke_XXXX.xcm)
cat > "$XCM" <<XCM_EOF
statistic chi
method leven 10 0.01
abund angr
xsect vern
cosmo 70 0 0.73
xset delta 0.01
systematic 0
lmod xsklineprofiles $PKG
model  kconv5*gaussian
          0.998      0.001     -0.998     -0.998      0.998      0.998
             60          1          0          0         90         90
              1        0.1          1          1        100        100
            400        0.1          1          1        400       1000
             50        0.1          1          1        400       1000
              3          1          0          0         10         10
              3          1          0          0      1e+10      1e+10
              3          1          0          0      1e+10      1e+10
              3          1          0          0      1e+10      1e+10
              3          1          0          0      1e+10      1e+10
              3          1          0          0      1e+10      1e+10
            6.5       0.05          0          0      1e+06      1e+06
            0.1       0.05          0          0         10         20
              1       0.01          0          0      1e+20      1e+24
bayes off
exit
XCM_EOF
OUT=$(xspec - "$XCM" 2>&1)
RC=$?
rm -f "$XCM"
echo "$OUT" | tail -30
if [ $RC -ne 0 ] || echo "$OUT" | grep -q "MODEL OUTPUT SET TO ZERO"; then
    echo "xspec_smoke: FAILED" >&2
    exit 1
fi
echo "xspec_smoke: ok"
