#!/bin/bash
# pin the oracle against the real PepQuery when a JVM and the built jar exist.
#   tools/java_reference_diff.sh /path/to/pepquery-2.0.4.jar [workdir]
# Writes the synthetic C1 inputs (MGF / FASTA / targets) and the oracle's tables, runs the reference on them for hyperscore,
# MVH and the unrestricted-modification search, and diffs psm_rank.txt / detail.txt / ptm.txt (tools/java_reference_diff.py).
# The jar is built from the unmodified reference with `mvn -q package` (needs the repositories of pom.xml:178-199).
set -e
JAR=${1:?usage: $0 pepquery-2.0.4.jar [workdir]}
W=${2:-/tmp/pq_java_diff}
command -v java >/dev/null || { echo "no java on PATH: parity stays unpinned (DESIGN.md §2)"; exit 2; }
cd "$(dirname "$0")/.."
rc=0
for cfg in "hs 1 0" "mvh 2 0" "fast 1 1"; do
  set -- $cfg; name=$1; method=$2; fast=$3
  python tools/java_reference_diff.py write "$W/$name" --method $method --fast $fast --step5
  extra=""; [ "$fast" = 1 ] && extra="-fast"
  java -jar "$JAR" -i "$W/$name/targets.txt" -ms "$W/$name/syn.mgf" -db "$W/$name/syn.fasta" -fixMod 1 -varMod 2 -n 1000 -m $method $extra \
       -cpu "$(nproc)" -o "$W/$name/java_out"
  python tools/java_reference_diff.py diff "$W/$name" "$W/$name/java_out" | tee "$W/$name/diff.log" || rc=1
done
exit $rc
