#!/bin/bash
# usage: scripts/sass_dump.sh <mangled-substring> > out.sass   (plain SASS of one kernel of libb2of.so)
SO=/root/repo/hackathonopticalflow_b200/csrc/libb2of.so
FN=$(cuobjdump -sass $SO | grep "Function :" | grep "$1" | head -1 | awk '{print $3}')
cuobjdump -sass -fun "$FN" $SO 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\/\*.*//'
