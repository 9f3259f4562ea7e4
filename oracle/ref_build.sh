#!/bin/bash
# TEST INFRASTRUCTURE -- compile the reference's OWN sources into oracle/_ref/ref_cli.
#
# Recipe: every libbreseq source listed in /root/reference/src/breseq/Makefile.am:38-71, compiled
# where it lies with the flags of Makefile.am:31-34 (-O3 -std=c++11), against
#   * oracle/_ref/config.h        hand-written stand-in for the autoconf header (version strings only)
#   * oracle/hts_shim/            htslib-compatible shim over zlib (htslib itself is absent) and
#                                 abort-stubs for the four miniz calls of the HTML-report zip writer
# and linked with oracle/ref_driver.cpp.  Reference sources are never copied: outputs (objects,
# config.h, the binary) go to oracle/_ref/ only, which is git-ignored but travels to the GPU box.
# The reference's own build system (autotools) is not run.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${BRESEQ_REFERENCE:-/root/reference}/src/breseq"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "ref_build: $REF not present; keeping any prebuilt $OUT/ref_cli"; exit 0; }
mkdir -p "$OUT/obj" "$OUT/share/breseq"
: > "$OUT/share/breseq/breseq_icon.png"   # Settings::set_global_paths() only tests that the data dir exists (settings.h:806-815)
cat > "$OUT/config.h" <<'EOF'
/* stand-in for the autoconf-generated config.h (configure.ac:26) */
#define PACKAGE_NAME "breseq"
#define PACKAGE_VERSION "0.50.0"
#define PACKAGE_STRING "breseq 0.50.0"
#define PACKAGE_BUGREPORT "jeffrey.e.barrick@gmail.com"
#define PACKAGE_URL "http://barricklab.org/breseq"
#define GITHUB_REVISION_STRING "oracle-build"
/* Makefile.am:33-34 passes this on the command line */
#define DATADIR "share/breseq/"
EOF
SRCS="alignment alignment_output anyoption calculate_trims candidate_junctions cn_evidence contingency_loci
coverage_output coverage_distribution dp_evidence error_count fasta fastq flagged_regions genome_diff genome_diff_entry
homologous_deletion identify_mutations mp_evidence pd_evidence mutation_predictor portable_random pgzstream pileup
output pileup_base reference_sequence resolve_alignments samtools_commands settings soft_clipping stats summary"
CXX="${CXX:-g++}"
FLAGS="-std=c++11 -O3 -w -I$OUT -I$HERE/hts_shim"
JOBS="${JOBS:-$(nproc)}"
stale=0
for s in $SRCS; do
  if [ ! -f "$OUT/obj/$s.o" ] || [ "$REF/$s.cpp" -nt "$OUT/obj/$s.o" ] || [ "$HERE/hts_shim/htslib/sam.h" -nt "$OUT/obj/$s.o" ]; then stale=1; fi
done
if [ $stale = 1 ]; then
  printf '%s\n' $SRCS | xargs -P "$JOBS" -I{} sh -c "$CXX $FLAGS -c $REF/{}.cpp -o $OUT/obj/{}.o"
fi
OBJS=""
for s in $SRCS; do OBJS="$OBJS $OUT/obj/$s.o"; done
$CXX $FLAGS -I"$REF" -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o"
$CXX -std=c++17 -O2 -w -I"$HERE/hts_shim" -c "$HERE/hts_shim/hts_shim.cpp" -o "$OUT/obj/hts_shim.o"
$CXX -o "$OUT/ref_cli" "$OUT/obj/ref_driver.o" $OBJS "$OUT/obj/hts_shim.o" -lz -lpthread
echo "ref_build: built $OUT/ref_cli"

# ---- the drop-in, compiled for real: the same reference objects with the two entry points' bodies renamed at compile time
# (-Derror_count=error_count_cpu -Didentify_mutations=identify_mutations_cpu: the reference's sources are compiled where they
# lie, nothing is copied) and adapters/breseq_adapter.cpp providing breseq::error_count() / breseq::identify_mutations() over
# libbrq.so.  ref_cli_brq takes ref_cli's command line; tests/test_gpu_adapter.py diffs what the two write.
REPO="$(cd "$HERE/.." && pwd)"
if [ -f "$REPO/breseq_b200/libbrq.so" ]; then
  RENAME="-Derror_count=error_count_cpu -Didentify_mutations=identify_mutations_cpu"
  for s in error_count identify_mutations; do
    if [ ! -f "$OUT/obj/${s}_renamed.o" ] || [ "$REF/$s.cpp" -nt "$OUT/obj/${s}_renamed.o" ]; then
      $CXX $FLAGS $RENAME -c "$REF/$s.cpp" -o "$OUT/obj/${s}_renamed.o"
    fi
  done
  $CXX $FLAGS -I"$REF" -I"$REPO/include" -c "$REPO/adapters/breseq_adapter.cpp" -o "$OUT/obj/breseq_adapter.o"
  OBJS_BRQ=""
  for s in $SRCS; do
    case $s in error_count|identify_mutations) OBJS_BRQ="$OBJS_BRQ $OUT/obj/${s}_renamed.o" ;; *) OBJS_BRQ="$OBJS_BRQ $OUT/obj/$s.o" ;; esac
  done
  $CXX -o "$OUT/ref_cli_brq" "$OUT/obj/ref_driver.o" "$OUT/obj/breseq_adapter.o" $OBJS_BRQ "$OUT/obj/hts_shim.o" \
       -L"$REPO/breseq_b200" -lbrq -Wl,-rpath,'$ORIGIN/../../breseq_b200' -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,/usr/local/cuda/lib64 -lz -lpthread
  echo "ref_build: built $OUT/ref_cli_brq"
fi
