#!/bin/bash
# TEST INFRASTRUCTURE ONLY. Builds the reference's own C++ library from the sources where they
# lie under /root/reference/cpp (unmodified, nothing copied) into oracle/_ref/ (git-ignored,
# travels to the GPU box with the gpurun snapshot).
#   libego.so          = the reference's libego (direct, acqmaxGP, maxRF, logCDFs)
#   libego_harness.so  = oracle/ref_harness.cpp + #include of the reference optimizeGP.cpp,
#                        plus direct.cpp, exposing per-candidate negei/negpi/negucb.
# The stock cpp/Makefile does not link on Linux; direct.cpp:575 needs
# pre-C++11 iostream semantics, hence -std=gnu++98 for the reference TUs.
set -e
REF=${IBO_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/cpp" ]; then
    echo "[build_ref] $REF/cpp not present; keeping prebuilt files in $OUT (if any)"; exit 0
fi
mkdir -p "$OUT/obj"
for f in direct optimizeGP optimizeRF helpers; do
    g++ -std=gnu++98 -O3 -fPIC -w -I"$REF/cpp" -c "$REF/cpp/$f.cpp" -o "$OUT/obj/$f.o"
done
g++ -shared -o "$OUT/libego.so" "$OUT/obj/direct.o" "$OUT/obj/optimizeGP.o" "$OUT/obj/optimizeRF.o" "$OUT/obj/helpers.o"
# harness: C++11 for std::thread; the included reference TU (optimizeGP.cpp) is C++11-clean.
g++ -std=gnu++11 -O3 -fPIC -w -pthread -I"$REF/cpp" -c "$HERE/ref_harness.cpp" -o "$OUT/obj/ref_harness.o"
g++ -shared -pthread -o "$OUT/libego_harness.so" "$OUT/obj/ref_harness.o" "$OUT/obj/direct.o"
echo "[build_ref] built $OUT/libego.so and $OUT/libego_harness.so"
