#!/bin/bash
# Final round-2 GPU run of the tree as committed: -m gpu suite, smoke, every bench line of profiles/README.md (one GPU).
mkdir -p gpurun_out
TAG=${1:-r02z}
python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
for w in skytem tempest mixed; do
  python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "$w rc=$?"
done
python - <<PY
import json
for w in ("", "_ref", "_skytem", "_tempest", "_mixed"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_bench%s.json" % w).read().strip().splitlines()[-1])
        print(w or "resolve", "%.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], d.get("clocks"))
    except Exception as e:
        print(w, "failed", e)
PY
