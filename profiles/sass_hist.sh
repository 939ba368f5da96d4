#!/bin/bash
# usage: profiles/sass_hist.sh OBJECT MANGLED_FUNCTION  -> opcode histogram of one kernel's SASS (and the listing in /tmp/sass.txt)
cuobjdump -sass -fun "$2" "$1" | grep -v "^\s*/\* 0x" > /tmp/sass.txt
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/sass.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{split($1,a,"."); print a[1]}' | sort | uniq -c | sort -rn
