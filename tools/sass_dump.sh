#!/bin/bash
# SASS of one kernel of a libhbk build: tools/sass_dump.sh <lib.so> <mangled-name substring> > out.sass
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/ {on = index($0, pat) > 0} on {print}'
