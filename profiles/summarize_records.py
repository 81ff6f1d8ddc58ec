"""Aggregate a BOFI_PROFILE_DUMP csv (per-launch CUDA-event records of the library) by kernel class and shape."""
import collections
import csv
import json
import sys

names = ['gemm_tcgen05', 'gemm_ffma', 'attention', 'layernorm', 'vocab_epilogue', 'other']
rows = list(csv.DictReader(open(sys.argv[1])))
agg = collections.OrderedDict()
for r in rows:
    k = (r['class'], r['d0'], r['d1'], r['d2'])
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += float(r['ms'])
    a[2] += float(r['gflop'])
tot = sum(v[1] for v in agg.values())
print('total kernel ms %.3f over %d launches' % (tot, len(rows)))
for k, (n, ms, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print('%-14s %-22s n=%3d ms=%7.3f per=%7.1fus share=%4.1f%% TF=%5.0f' % (names[int(k[0])], 'x'.join(k[1:]), n, ms, ms / n * 1e3, 100 * ms / tot, gf / ms if ms else 0))
if len(sys.argv) > 3:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print('bench: %.0f captions/s, %.2f ms/step, e2e %.0f, tcgen05 GEMM %.0f TFLOP/s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved']))
