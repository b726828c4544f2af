#!/bin/bash
# Round profile capture (run under gpurun, ONE GPU): launch list of the bench command's timed steps + one full capture of every hot
# kernel, summarised to text ON THE BOX (gpurun only copies 64 MiB back). Outputs: gpurun_out/<tag>_*.txt|csv (+ the .ncu-rep files
# while they fit). usage: tools/profile_round.sh <tag> [worlds]
TAG=${1:-r1}; WORLDS=${2:-256}
export B2J_BENCH_CUPROFILE=1 B2J_BATCH_GROUPS=1
mkdir -p gpurun_out
HOT='KFindPairs|KProcessPairs|KCopyCached|KCollideConvex|KCollideEpa|KFinishPairs|KSetupConstraints|KSolveVelocity'
for PHASE in impact:60 rest:110; do
  NAME=${PHASE%%:*}; WARM=${PHASE##*:}
  BENCH="python bench.py --worlds $WORLDS --steps 1 --warmup $WARM --no-cpu-baseline --no-pile"
  # every launch of the timed step with its device time
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_${NAME}.csv $BENCH > gpurun_out/${TAG}_launches_${NAME}.log 2>&1
  python tools/ncu_launches.py gpurun_out/${TAG}_launches_${NAME}.csv > gpurun_out/${TAG}_launches_${NAME}.txt 2>&1
  # full metric set: the first launch of each hot kernel of the step (+ the first velocity phases)
  ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:"$HOT" -c 12 -f -o gpurun_out/${TAG}_hot_${NAME} $BENCH > gpurun_out/${TAG}_hot_${NAME}.log 2>&1
  python tools/ncu_summary.py gpurun_out/${TAG}_hot_${NAME}.ncu-rep > gpurun_out/${TAG}_hot_${NAME}.txt 2>&1
done
du -sh gpurun_out; ls -la gpurun_out | grep ${TAG}_
# stay below the copy-back limit: drop the binary reports first if needed
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/${TAG}_hot_impact.ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/${TAG}_hot_rest.ncu-rep; fi
