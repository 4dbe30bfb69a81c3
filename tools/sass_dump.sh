#!/bin/bash
# developer aid: nvdisasm -g of one kernel of a library -> stdout.  usage: tools/sass_dump.sh <lib.so> <kernel substring>
L=$(realpath "$1"); T=$(mktemp -d); cd $T; cuobjdump -xelf all "$L" >/dev/null 2>&1
for f in *.cubin; do nvdisasm -g $f 2>/dev/null | awk -v k="$2" '/^\t\.section\t\.text\./{f=index($0,k)>0} /^\t\.section/&&!/\.text\./{f=0} f'; done
rm -rf $T
