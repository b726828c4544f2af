#!/bin/bash
# Round profile capture (run under gpurun, ONE GPU): launch list of one timed step of the bench command + one full capture of every hot
# kernel, summarised to text ON THE BOX (gpurun only copies 64 MiB back). Outputs: gpurun_out/<tag>_*.txt|csv|json (+ the .ncu-rep
# files while they fit). usage: tools/profile_round.sh <tag> [worlds]
# Steps: "impact" = step 22 of the driver's bench window (steps 5..25; the boxes land in steps 19..24: GJK / EPA heavy), "rest" = step 110
# (pyramids at rest). Programmatic dependent launch is off under ncu (it serialises kernels anyway).
TAG=${1:-r2}; WORLDS=${2:-256}
export B2J_BENCH_CUPROFILE=1 B2J_BATCH_GROUPS=1 B2J_SOLVE_PDL=0
mkdir -p gpurun_out
HOT='KFindPairs|KProcessPairs|KCopyCached|KCollideConvex|KCollideEpa|KFinishPairs|KSetupConstraints|KSolveVelocity|KSolvePosition|solve_velocity_tma|sched_block'
for PHASE in impact:22 rest:110; do
  NAME=${PHASE%%:*}; WARM=${PHASE##*:}
  BENCH="python bench.py --worlds $WORLDS --steps 1 --warmup $WARM --no-cpu-baseline --no-pile --no-extras"
  # every launch of the timed step with its device time
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches_${NAME}.csv $BENCH > gpurun_out/${TAG}_launches_${NAME}.log 2>&1
  python tools/ncu_launches.py gpurun_out/${TAG}_launches_${NAME}.csv > gpurun_out/${TAG}_launches_${NAME}.txt 2>&1
  # full metric set: the first launches of each hot kernel of the step
  ncu --set full --clock-control none --profile-from-start off --import-source on --kernel-name-base demangled -k regex:"$HOT" -c 14 -f -o gpurun_out/${TAG}_hot_${NAME} $BENCH > gpurun_out/${TAG}_hot_${NAME}.log 2>&1
  python tools/ncu_summary.py gpurun_out/${TAG}_hot_${NAME}.ncu-rep > gpurun_out/${TAG}_hot_${NAME}.txt 2>&1
done
# DRAM traffic of the velocity solve kernel against its algorithmic bytes (bench.py's roofline.traffic)
python tools/ncu_traffic.py gpurun_out/${TAG}_hot_rest.ncu-rep gpurun_out/${TAG}_hot_rest.log > gpurun_out/${TAG}_solve_traffic.json 2> gpurun_out/${TAG}_solve_traffic.err
# the one launch TMA solve (B2J_SOLVE_MODE=2) on the same resting step
B2J_SOLVE_MODE=2 ncu --set full --clock-control none --profile-from-start off --import-source on --kernel-name-base demangled -k regex:"solve_velocity_tma|solve_position_all" -c 2 -f -o gpurun_out/${TAG}_hot_tma \
  python bench.py --worlds $WORLDS --steps 1 --warmup 110 --no-cpu-baseline --no-pile --no-extras > gpurun_out/${TAG}_hot_tma.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_hot_tma.ncu-rep > gpurun_out/${TAG}_hot_tma.txt 2>&1
# the warp cooperative mesh kernel on the ConvexVsMesh scene (single world, steps 0..260: the bodies land on the terrain)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"run_kernel_warp_coop" -s 100 -c 3 -f -o gpurun_out/${TAG}_hot_mesh python tools/diag_mesh.py > gpurun_out/${TAG}_hot_mesh.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_hot_mesh.ncu-rep > gpurun_out/${TAG}_hot_mesh.txt 2>&1
du -sh gpurun_out; ls -la gpurun_out | grep ${TAG}_
# stay below the copy-back limit: drop the binary reports first if needed
for f in hot_impact hot_rest hot_tma hot_mesh; do if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/${TAG}_$f.ncu-rep; fi; done
