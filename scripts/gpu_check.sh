#!/bin/bash
# Regular GPU check: all gpu tests, smoke, bench (device numbers + stage table).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-check}
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "tflops", round(d["tflops"],1), "clocks", d["clocks"])
for k, v in (d.get("stages") or {}).items():
    print(f"  {k:8s} {v['ms_per_event']*1e3:8.1f} us  " + (f"{v['tflops']:7.1f} TF/s {100*v['frac_of_bf16_peak']:5.1f}%" if 'tflops' in v else f"{v['gbs']:7.1f} GB/s {100*v['frac_of_hbm_peak']:5.1f}%"))
print("cpu", d.get("cpu_baseline"))
PY
