#!/bin/bash
# usage: scripts/sass_kernel.sh <object.o> <mangled-kernel-substring> <out.txt>
# nvdisasm listing (with source-line markers) of ONE kernel of an object file; prints its instruction count.
set -e
obj=$(realpath "$1"); pat=$2; out=$(realpath -m "$3")
tmp=$(mktemp -d)
(cd "$tmp" && cuobjdump -xelf all "$obj" > /dev/null && nvdisasm -g -c "$tmp"/*.cubin > "$tmp/all.txt")
s=$(grep -n "^\.text\..*${pat}" "$tmp/all.txt" | head -1 | cut -d: -f1)
e=$(awk -v s="$s" 'NR>s && /^\/\/-+ \.text/ {print NR; exit}' "$tmp/all.txt")
[ -z "$e" ] && e=$(wc -l < "$tmp/all.txt")
sed -n "${s},${e}p" "$tmp/all.txt" | grep -v 'inlined at\|^$' | grep -v '//## File "/usr' \
  | sed -e "s|File \".*/csrc/|File \"|" > "$out"
echo "$(grep -c '^\s*/\*[0-9a-f]*\*/' "$out") SASS instructions -> $out"
rm -r "$tmp"
