#!/bin/bash
# usage: tools/poll.sh <marker-file-in-gpurun_out> ; prints DONE + tails when present
f=/root/repo/gpurun_out/$1
if [ -f "$f" ]; then echo DONE; else echo waiting; /usr/local/graft/bin/gpurun --status 2>/dev/null | grep -E "running|elapsed|gpu_minutes_left" | head -5; fi
