#!/bin/bash
# All N-GPU measurements of one gpurun --gpus N call.  usage: tools/mg_run.sh N
# (bench.py under torchrun for config 2 (weak, views) and config 4 (strong, ray tiles); the one-process C++ driver
#  build/volrend_headless_mg on both trees with --check; config 5 (8 scenes) through tools/run_configs.py)
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r02_mg_$N
PORT=29517
run_bench() {  # $1 = extra args, $2 = output name
  if [ "$N" = "1" ]; then timeout 600 python bench.py $1 > $OUT.$2.json 2> $OUT.$2.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N $1 > $OUT.$2.json 2> $OUT.$2.err; fi
  tail -c 300 $OUT.$2.err | grep -i "error\|Traceback" | head -3
}
run_bench "--no-cli" config2
run_bench "--workload config4" config4
# one process, N devices: the C++ driver
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from volrend_b200 import synth
import numpy as np
if not os.path.exists("/tmp/vr_bench_tree.npz"):
    synth.make_tree("lego", depth=10, basis_dim=16, seed=0).save_npz("/tmp/vr_bench_tree.npz")
if not os.path.exists("/tmp/vr_config4_tree.npz"):
    synth.make_tree("gyroid_small", depth=11, basis_dim=25, seed=0, band_cells=1.0).save_npz("/tmp/vr_config4_tree.npz")
synth.write_pose_files(synth.nerf_synthetic_test_poses(200), "/tmp/vr_poses200", synth.focal_for(800))
synth.write_pose_files(synth.nerf_synthetic_test_poses(40, radius=1.6, elev_deg=25.0), "/tmp/vr_poses40", 1500.0)
PY
MG=build/volrend_headless_mg
if [ -x $MG ]; then
  { echo "== config 2 tree, 200 poses 800x800, views mode, $N GPU(s)"
    timeout 300 $MG /tmp/vr_bench_tree.npz -w 800 -h 800 --fx 1111.111 --gpus $N --mode views --reps 3 --check /tmp/vr_poses200/pose/*.txt 2>&1 | grep -E "per frame|fps|Mrays|check|rror"
    echo "== config 2 tree, tiles mode"
    timeout 300 $MG /tmp/vr_bench_tree.npz -w 800 -h 800 --fx 1111.111 --gpus $N --mode tiles --reps 3 --check /tmp/vr_poses200/pose/*.txt 2>&1 | grep -E "per frame|fps|Mrays|check|rror"
    echo "== config 4 tree, 40 poses 1920x1080, tiles mode"
    timeout 300 $MG /tmp/vr_config4_tree.npz -w 1920 -h 1080 --fx 1500 --gpus $N --mode tiles --reps 3 --check /tmp/vr_poses40/pose/*.txt 2>&1 | grep -E "per frame|fps|Mrays|check|rror"
    echo "== config 4 tree, tiles mode, one frame per launch (--batch 1)"
    timeout 300 $MG /tmp/vr_config4_tree.npz -w 1920 -h 1080 --fx 1500 --gpus $N --mode tiles --batch 1 --reps 2 /tmp/vr_poses40/pose/*.txt 2>&1 | grep -E "per frame|fps|Mrays|check|rror"
  } > $OUT.mg_cli.log 2>&1
fi
# config 5: 8 scenes, one (or 8/N) per GPU
if [ "$N" = "1" ]; then timeout 900 python tools/run_configs.py 5 > $OUT.config5.log 2>&1
else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) tools/run_configs.py 5 > $OUT.config5.log 2>&1; fi
grep -h "^config5" $OUT.config5.log | cut -c1-400
for f in config2 config4; do python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$OUT.$f.json") if l.startswith("{")][-1]
    print("$f N=$N value %.1f ms/step %.3f e2e %.1f frac %.3f kernel_ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"]), d["config"].get("reassembly_identical_to_single_gpu"))
except Exception as e:
    print("$f N=$N: no line", e)
PY
done
cat $OUT.mg_cli.log
