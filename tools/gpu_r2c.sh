#!/bin/bash
# round 2, GPU call C: what the bounding loop costs (1 step instead of 19), LayerNorm fused into the small GEMMs, attention source page
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_selfcritical.py -m gpu -q --timeout 900 -rA > $O/c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c_pytest.log
timeout 600 python bench.py --no-extras > $O/c_base.json 2> $O/c_bench.err
BOFI_DEBUG_MAX_BOUND_STEPS=1 timeout 600 python bench.py --no-extras > $O/c_bound1.json 2>> $O/c_bench.err
timeout 600 python bench.py --no-extras --depth 1 > $O/c_base_d1.json 2>> $O/c_bench.err
BOFI_DEBUG_MAX_BOUND_STEPS=1 timeout 600 python bench.py --no-extras --depth 1 > $O/c_bound1_d1.json 2>> $O/c_bench.err
BOFI_LNFUSE_SMALL=1 timeout 600 python bench.py --no-extras > $O/c_lnsmall.json 2>> $O/c_bench.err
BOFI_LNFUSE_SMALL=1 timeout 600 python bench.py --no-extras --depth 1 > $O/c_lnsmall_d1.json 2>> $O/c_bench.err
BOFI_LNFUSE_SMALL=1 timeout 300 python bench.py --batch 1 --depth 1 --calib s_cap --no-extras --steps 50 > $O/c_lnsmall_lat1.json 2>> $O/c_bench.err
BOFI_GRAPH=0 timeout 300 python tools/one_decode.py > $O/c_one_decode.log 2>&1 && \
BOFI_GRAPH=0 timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attention_mma_kernel -s 2 -c 1 -o /tmp/full_att python tools/one_decode.py > $O/c_ncu_att.log 2>&1
ncu -i /tmp/full_att.ncu-rep --page source --csv > /tmp/att_source.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/att_source.csv')))
hdr = rows[0]
print(hdr[:12])
keep = [i for i, h in enumerate(hdr) if h in ('#', 'Address', 'Source', 'Warp Stall Sampling (All Samples)', 'Warp Stall Sampling (Not-issued Samples)', 'Instructions Executed', '# Samples')]
samp = [i for i, h in enumerate(hdr) if 'Sampling (All' in h]
body = rows[1:]
def val(r):
    try:
        return float(r[samp[0]])
    except Exception:
        return 0.0
body.sort(key=val, reverse=True)
with open('gpurun_out/c_att_source_top.csv', 'w') as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in keep])
    for r in body[:120]:
        w.writerow([r[i] for i in keep])
PY
du -sh $O
