# ASan + UBSan builds of the oracle and of libwpt's host code (builders, OBJ parser, ABI guards, host-only sessions), then the
# whole CPU test suite against them. Runs without a GPU.
# usage: bash scripts/host_sanitize.sh > profiles/r2_host_sanitizers.log 2>&1
set -e
D=${TMPDIR:-/tmp}/wpt_san; mkdir -p $D
SAN="-fsanitize=address,undefined -fno-sanitize-recover=undefined"
(cd oracle && g++ -std=c++17 -O1 -g -fPIC -ffp-contract=off -fno-fast-math -pthread $SAN -shared -o $D/liboracle_asan.so oracle_capi.cpp)
(cd wasm_pathtracer_b200/csrc && cp ptxas.log $D/ptxas.keep 2>/dev/null; make -s OUT=$D/libwpt_asan.so EXTRA="-Xcompiler -fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=undefined,-g"; cp $D/ptxas.keep ptxas.log 2>/dev/null || true)
echo "== CPU suite on the ASan + UBSan builds (any report aborts the run)"
WPT_LIBRARY=$D/libwpt_asan.so WPT_ORACLE_LIBRARY=$D/liboracle_asan.so \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 \
python -m pytest tests -x -q -m "not gpu" -p no:cacheprovider
echo "== exit code $?"
