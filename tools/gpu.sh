#!/bin/bash
# build the library, then run a command on a B200 box:  tools/gpu.sh [--gpus N] <timeout-s> '<command>'
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
make -C /root/repo/mellon_b200/csrc -j8 > /tmp/mb_make.log 2>&1 || { grep -E "error" -A3 /tmp/mb_make.log | head -40; exit 9; }
(cd /root/repo && python -c "from mellon_b200 import _native as n; n.load_library()") || exit 9
cd /root/repo && exec /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@"
