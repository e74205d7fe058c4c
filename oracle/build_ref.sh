#!/bin/bash
# TEST INFRASTRUCTURE.  Compiles the reference's own solve-phase translation units, unmodified
# and in place under /root/reference/src, against oracle/ref_shim (hypre stub + prelude), and
# links them with oracle/ref_driver.cpp into oracle/_ref/libref_smem.so.  Flags are the
# reference's own: g++ -fopenmp -O3 (/root/reference/Makefile:37).  Outputs only into
# oracle/_ref/ (git-ignored; travels to the GPU box).  The reference's build system is not run.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference/src
OUT="$HERE/_ref"
mkdir -p "$OUT"
CXX="g++ -O3 -fopenmp -fPIC -w -I$HERE/ref_shim -I$REF -include $HERE/ref_shim/ref_prelude.hpp"
for f in SMEM_MatVec SMEM_Smooth SMEM_Sync_AMG SMEM_Async_AMG SMEM_ExtendedSystem SMEM_Cheby SEQ_MatVec SEQ_Smooth SEQ_AMG Misc DMEM_Mult DMEM_Misc DMEM_Add DMEM_Smooth; do
   $CXX -c "$REF/$f.cpp" -o "$OUT/$f.o"
done
# SMEM_Setup.cpp (SmoothTransfer, the thread partition, the work model): Eigen is un-vendored -- oracle/ref_shim/eigen_stub stands in
$CXX -I"$HERE/ref_shim/eigen_stub" -c "$REF/SMEM_Setup.cpp" -o "$OUT/SMEM_Setup.o"
# BuildHypreMatrix.cpp (stencil coefficients of -problem 7pt / 27pt / difconv): its own header is missing from the reference; same stand-in
$CXX -I"$HERE/ref_shim/eigen_stub" -c "$REF/BuildHypreMatrix.cpp" -o "$OUT/BuildHypreMatrix.o"
# SMEM_Solve.cpp: its printf (residual history, src/SMEM_Solve.cpp:95-103,232-239) goes to the hook
$CXX -c "$HERE/ref_shim/wrap_SMEM_Solve.cpp" -o "$OUT/SMEM_Solve.o"
$CXX -c "$HERE/ref_driver.cpp" -o "$OUT/ref_driver.o"
g++ -shared -fopenmp -o "$OUT/libref_smem.so" "$OUT"/*.o -Wl,--no-undefined 2> "$OUT/link.log" || {
   # report what the stubs still miss, then link lazily (unreached symbols stay unresolved)
   grep -o "undefined reference to \`[^']*'" "$OUT/link.log" | sort -u | head -40
   g++ -shared -fopenmp -o "$OUT/libref_smem.so" "$OUT"/*.o
}
# the same objects + the binding of INTEGRATION.md (integration/SMEM_B200.hpp compiled against the reference's Main.hpp), linked
# against the product library: the reference-side structs drive libamg_b200.so (GPU test only; needs the built product library)
PKG="$HERE/../async-multigrid_b200"
if [ -f "$PKG/libamg_b200.so" ]; then
   $CXX -DREF_WITH_B200 -I"$HERE/../include" -I"$HERE/../integration" -c "$HERE/ref_driver.cpp" -o "$OUT/ref_driver.o"
   g++ -shared -fopenmp -o "$OUT/libref_b200.so" "$OUT"/*.o -L"$PKG" -lamg_b200 -Wl,-rpath,'$ORIGIN/../../async-multigrid_b200' \
      && echo "built $OUT/libref_b200.so"
fi
rm -f "$OUT"/*.o
echo "built $OUT/libref_smem.so"
