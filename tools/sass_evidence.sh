#!/bin/bash
# Writes profiles/r02_sass_{fq_mul,fr_mul,fr_lazy,k_accumulate,k_aff_finish}.txt: SASS listings (cuobjdump) of the field
# primitives (tools/sass_probe.cu) and of the two multiplier-bound MSM kernels of the shipped library, each headed by its
# opcode histogram.  Runs on the CPU box (nvcc cross-compiles, cuobjdump needs no GPU).
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles; B=$(mktemp -d)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -cubin -o $B/sass_probe.cubin tools/sass_probe.cu
hist() { grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn; }
one() {  # $1 = file with the sass of all functions, $2 = function name pattern, $3 = output
  awk -v pat="$2" '/Function :/{f = ($0 ~ pat)} f' "$1" > $B/_one.sass
  { echo "# $2  (cuobjdump -sass, sm_100a) - opcode histogram, then the listing"; hist < $B/_one.sass | head -40; echo; cat $B/_one.sass; } > "$3"
}
cuobjdump -sass $B/sass_probe.cubin > $B/sass_probe.sass
one $B/sass_probe.sass probe_fq_mul profiles/r02_sass_fq_mul.txt
one $B/sass_probe.sass probe_fr_mul profiles/r02_sass_fr_mul.txt
one $B/sass_probe.sass probe_fr_lazy profiles/r02_sass_fr_lazy_mul_add.txt
cuobjdump -sass gemini_b200/libgemini_b200.so > $B/lib.sass
one $B/lib.sass 'k_accumulateILb1' $B/_acc.txt
one $B/lib.sass 'k_aff_finishILb0' $B/_fin.txt
# the full listings of the big kernels are 15k lines each: keep the histogram and the first 400 lines
head -450 $B/_acc.txt > profiles/r02_sass_k_accumulate.txt
head -450 $B/_fin.txt > profiles/r02_sass_k_aff_finish.txt
for f in profiles/r02_sass_*.txt; do echo "$f: $(head -12 $f | tail -11 | tr '\n' ';')"; done
