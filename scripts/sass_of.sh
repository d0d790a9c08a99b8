#!/bin/bash
# usage: scripts/sass_of.sh <object> <mangled-name regex>  -> one SASS instruction per line
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function : /{f = ($0 ~ pat)} f' | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed 's/\/\* 0x[0-9a-f]* \*\///' | awk '{$1=$1};1'
