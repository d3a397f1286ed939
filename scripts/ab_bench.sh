#!/bin/bash
# Same-call A/B of library builds (run under gpurun): ms/step differs by +-3 % between GPU boxes, so two kernel variants are
# only comparable when they run back to back on the same box.  Put the builds to compare in _ab/<name>.so (git-ignored; copy
# melissa_b200/lib/libmelissa_b200.so there after building each variant) and run
#     gpurun -- 'bash scripts/ab_bench.sh base variant1 variant2'
# The last library listed stays installed; rebuild afterwards (python -m melissa_b200.build --force).
B="python bench.py --no-e2e --no-cpu-baseline --no-train --no-extra --no-flip --steps 20 --warmup 5"
for i in 1 2; do for v in "$@"; do
  cp _ab/$v.so melissa_b200/lib/libmelissa_b200.so
  timeout 300 $B > gpurun_out/ab_${v}_$i.json 2> gpurun_out/ab_${v}_$i.err || echo "$v run $i failed"
done; done
python - "$@" <<'PY'
import json, sys
for i in (1, 2):
    for v in sys.argv[1:]:
        try:
            d = json.loads(open(f"gpurun_out/ab_{v}_{i}.json").read().strip().splitlines()[-1])
            print(f"{v:12s} {d['value'] / 1e6:8.2f} M/s {d['ms_per_step']:.4f} ms/step  conv2 {d['roofline']['kernel_ms']:.3f}  conv1 stage "
                  f"{d['roofline_conv1']['kernel_ms']:.3f}  proj2 {d['roofline_tensor']['kernel_ms']:.3f}")
        except Exception as e:
            print(v, "no result:", e)
PY
