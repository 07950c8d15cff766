#!/bin/bash
# usage: dev/sass_of.sh <lib.so> <substring of mangled kernel name> -> clean SASS listing on stdout
lib=$1; pat=$2
cuobjdump -sass "$lib" | awk -v pat="$pat" '
/Function :/ { on = (index($0, pat) > 0) }
on && /^ +\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\// { sub(/\/\* 0x[0-9a-f]+ \*\//, ""); print }'
