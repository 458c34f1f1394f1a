#!/bin/bash
# Round-end style check on one box: all gpu tests, smoke, the driver's bench command and its reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-final}
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
S=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $? in $(( $(date +%s) - S )) s"; tail -2 gpurun_out/${TAG}_bench.err
S=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm exit $? in $(( $(date +%s) - S )) s"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "fwd ms", d["e2e"].get("forward_ms_in_pipeline"), "clocks", d["clocks"])
print("roofline", d["roofline"]); print("parity", d["parity"])
for k, v in (d.get("stages") or {}).items():
    print(f"  {k:8s} {v['ms_per_event']*1e3:8.1f} us  " + (f"{v['tflops']:7.1f} TF/s {100*v['frac_of_bf16_peak']:5.1f}%" if 'tflops' in v else f"{v['gbs']:7.1f} GB/s {100*v['frac_of_hbm_peak']:5.1f}%"))
r = json.loads(open("gpurun_out/${TAG}_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", r.get("value"), r.get("e2e"), r.get("cpu_baseline", {}).get("kind"))
PY
