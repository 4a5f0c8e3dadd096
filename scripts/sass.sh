#!/bin/bash
# Dump the SASS of one kernel of libmovfe.so: scripts/sass.sh <unit: extract|grid|pose|raster> <kernel regex> > out.sass
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
cd "$T"
cuobjdump -xelf "$1" "$ROOT/mov-slam_b200/lib/libmovfe.so" > /dev/null
nvdisasm -c "$1"*.cubin | awk -v k="$2" '/^\.text\./{f = ($0 ~ k)} f'
