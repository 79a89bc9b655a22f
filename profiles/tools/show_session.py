"""Print the recommend sweep / cfg5 / recommend launch list of the last gpurun session (gpurun_out/)."""
import csv, json, os, re, sys
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "gpurun_out")
for l in open(os.path.join(R, "recommend_sweep.jsonl")):
    d = json.loads(l); v = d['variant']
    print(v, 'total %.2f gemm %.2f TF %.0f (%.3f) users/s %.2fM overlap %.3f redo %s' % (d['ms_total'], d['ms_gemm_filter'], d['tflops_gemm_filter'], d['frac_of_bf16_peak'], d['users_per_s'] / 1e6, d['topk_overlap_vs_exact'], d['rows_redone_on_exact_path']))
try:
    d = json.load(open(os.path.join(R, "bench_cfg5.json"))); r = d['recommend']
    print('cfg5 total %.1f gemm %.1f TF %.0f frac %.3f users/s %.2fM ov %.3f redo %s' % (r['ms_total'], r['ms_gemm_filter'], r['tflops_gemm_filter'], r['frac_of_bf16_peak'], r['users_per_s'] / 1e6, r['topk_overlap_vs_exact'], r['rows_redone_on_exact_path']), d['clocks'])
except Exception as e:
    print("cfg5:", e)
rows = list(csv.reader(open(os.path.join(R, "launches_recommend.csv"))))
for k, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr = r; start = k; break
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
names = [(re.sub(r'\(.*', '', r[ki])[:70], float(r[vi].replace(',', '')) / 1e6) for r in rows[start + 1:] if len(r) > vi]
first = next(k for k, (n, _) in enumerate(names) if 'pack_gemm_users' in n)
for n, ms in names[first:first + 6]:
    print("%-70s %10.3f ms" % (n, ms))
