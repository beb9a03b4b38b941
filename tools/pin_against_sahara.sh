#!/usr/bin/env bash
# pin_against_sahara.sh — pins this repo against a REAL sahara binary (seqan/sahara built with its CPM dependencies,
# which this container cannot do: DESIGN.md §2, "parity unpinned").  Run it on a machine that has both; it is the check
# SURVEY.md §8c asks the build to ship.  Nothing in tests/, bench.py or smoke() depends on it.
#
#   tools/pin_against_sahara.sh /path/to/sahara genome.fasta reads.fasta [k=2] [generator=h2-k2]
#
# Steps (every one prints PASS / FAIL; exit code = number of failures):
#   1. index file: `sahara index` (src/sahara/index.cpp:87-100) vs `sahara_b200/sahara index` — byte compare
#   2. scheme tables: `sahara search_scheme -a --columba DIR -k K` (src/sahara/search_scheme.cpp:252-276) dumps the
#      upstream pi/L/U tables; ours are printed by the Python mirror and compared as text
#   3. hits: `sahara search` (src/sahara/search.cpp:104-274) vs ours on the REAL index file and the REAL tables
#      (--scheme-file), edit distance and Hamming, compared as `sort -u` of the "queryId seqId pos" lines and as
#      plain `sort` (multiplicities)
#   4. --max_hits 1 / 5 (search_n, src/sahara/search.cpp:228,231) and -m besthits (src/sahara/search.cpp:233-240),
#      compared as sorted lines — these depend on the recursion order reconstructed in SURVEY.md §9.4
set -u
SAHARA=${1:?path to the reference sahara binary}
GENOME=${2:?genome fasta}
READS=${3:?reads fasta}
K=${4:-2}
GEN=${5:-h2-k2}
HERE=$(cd "$(dirname "$0")/.." && pwd)
OURS="$HERE/sahara_b200/sahara"
WORK=$(mktemp -d)
fails=0
verdict() { if [ "$1" -eq 0 ]; then echo "PASS  $2"; else echo "FAIL  $2"; fails=$((fails + 1)); fi; }

# 1. index file
cp "$GENOME" "$WORK/ref.fa"; cp "$GENOME" "$WORK/ours.fa"
"$SAHARA" index "$WORK/ref.fa" > "$WORK/ref_index.log" 2>&1 || { echo "reference sahara index failed"; cat "$WORK/ref_index.log"; exit 99; }
"$OURS" index "$WORK/ours.fa" > "$WORK/ours_index.log" 2>&1
cmp -s "$WORK/ref.fa.idx" "$WORK/ours.fa.idx"; verdict $? "index file byte-identical (ref.fa.idx vs ours.fa.idx)"

# 2. scheme tables
"$SAHARA" search_scheme -a --columba "$WORK/columba" -k "$K" > /dev/null 2>&1
TABLE="$WORK/columba/$GEN/$K/searches.txt"
if [ -f "$TABLE" ]; then
    PYTHONPATH="$HERE" python - "$GEN" "$K" > "$WORK/ours_searches.txt" <<'EOF'
import sys
import sahara_b200 as sb
print(sb.SearchScheme.generate(sys.argv[1], 0, int(sys.argv[2])).to_columba(), end="")
EOF
    diff -q <(sort "$TABLE") <(sort "$WORK/ours_searches.txt") > /dev/null; verdict $? "scheme table of $GEN, k=$K equals upstream's"
    SCHEME=(--scheme-file "$TABLE")
else
    echo "SKIP  upstream dumped no table for $GEN k=$K"; SCHEME=()
fi

# 3. + 4. hits
run_pair() {  # name, extra flags...
    local name=$1; shift
    "$SAHARA" search -q "$READS" -i "$WORK/ref.fa.idx" -e "$K" -g "$GEN" -o "$WORK/ref_$name.txt" "$@" > "$WORK/ref_$name.log" 2>&1
    "$OURS" search -q "$READS" -i "$WORK/ref.fa.idx" -e "$K" -g "$GEN" -o "$WORK/ours_$name.txt" "${SCHEME[@]}" "$@" > "$WORK/ours_$name.log" 2>&1
    local rc=$?
    [ $rc -eq 0 ] || { verdict 1 "$name: our search failed (see $WORK/ours_$name.log)"; return; }
    diff -q <(sort -u "$WORK/ref_$name.txt") <(sort -u "$WORK/ours_$name.txt") > /dev/null; verdict $? "$name: hit set (sort -u)"
    diff -q <(sort "$WORK/ref_$name.txt") <(sort "$WORK/ours_$name.txt") > /dev/null; verdict $? "$name: hit multiset (sort)"
}
run_pair lev -d lev
run_pair ham -d ham
run_pair lev_max1 -d lev --max_hits 1
run_pair lev_max5 -d lev --max_hits 5
run_pair ham_max1 -d ham --max_hits 1
# besthits generates its schemes per stratum (generator(j, j)): the upstream tables of the strata are not in one file
SCHEME=()
run_pair besthits -m besthits
run_pair besthits_max2 -m besthits --max_hits 2

echo "work directory: $WORK"
exit $fails
