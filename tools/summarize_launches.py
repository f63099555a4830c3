"""profiles/launches_rNN_summary.md from an ncu launch list CSV (gpu__time_duration.sum per launch)."""
import csv, collections, sys
src = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_r01.csv'
dst = sys.argv[2] if len(sys.argv) > 2 else 'profiles/launches_r01_summary.md'
with open(src) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def dur(row):
    v = float(row['Metric Value']); u = row['Metric Unit']
    return v / 1000.0 if u == 'ns' else (v * 1000 if u == 'ms' else v)
def short(n):
    return n.split('(')[0].replace('void gitb200::', '').replace('gitb200::', '').replace('void ', '')
names = [short(r['Kernel Name']) for r in rows]
emb = [i for i, n in enumerate(names) if n.startswith('embed_ln')]
enc_end = emb[0]
fa = [i for i in range(enc_end) if names[i].startswith('flash_attn')]
n_enc_layers = len(fa) - 5          # the prefill runs attention in 5 of its 6 layers
i = fa[n_enc_layers - 1] + 1; cnt = 0
while cnt < 3:
    if names[i].startswith('gemm'): cnt += 1
    i += 1
enc_stop = i + 1
def agg(lo, hi):
    a = collections.OrderedDict(); tot = 0
    for i in range(lo, hi):
        k = names[i] + ' grid=' + rows[i]['Grid Size']
        x = a.setdefault(k, [0, 0.0]); x[0] += 1; x[1] += dur(rows[i]); tot += dur(rows[i])
    return a, tot
out = []
a, tenc = agg(0, enc_stop); out.append(('encoder (patch embed + ViT blocks + ln_post)', a, tenc))
a, tpre = agg(enc_stop, enc_end); out.append(('prefill (visual projection + image rows of the 6 decoder layers, fills the image K/V cache)', a, tpre))
a, tdec = agg(emb[1], emb[2]); out.append(('one decode step (2nd captured; 39 per caption batch)', a, tdec))
txt = ['# ncu launch list, round 1 (cold-cache, serialised per-launch times: compare SHARES, not absolutes)\n',
       'command: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 420 --csv python bench.py --steps 1 --warmup 3 --pipeline 1 --no-cpu-baseline --no-micro --ncu-range\n',
       '(profiler range = the timed step, one batch at a time; the first 420 launches = encoder + prefill + first decode steps; raw list: launches_r01.csv)\n\n']
whole = tenc + tpre + 39 * tdec
for title, a, t in out:
    txt.append('## %s: %.1f us total (%d launches)\n' % (title, t, sum(x[0] for x in a.values())))
    for k, x in sorted(a.items(), key=lambda kv: -kv[1][1]):
        txt.append('  %-64s n=%3d total=%8.1f us avg=%7.2f us  %5.1f%%\n' % (k[:64], x[0], x[1], x[1] / x[0], 100 * x[1] / t))
    txt.append('\n')
txt.append('## whole step under ncu: encoder %.0f + prefill %.0f + 39 x %.0f decode = %.0f us\n' % (tenc, tpre, tdec, whole))
gem = sum(x[1] for title, a, t in out[:2] for k, x in a.items() if k.startswith('gemm'))
gemdec = sum(x[1] for k, x in out[2][1].items() if k.startswith('gemm'))
txt.append('gemm{,2}_bf16_tcgen05 share: encoder+prefill %.0f us + decode 39 x %.0f us = %.1f%% of the step (bench.py in situ: encoder+prefill ~4.4 ms + 39 x ~0.34 ms decode = 17.9 ms per batch, GEMM kernels ~55%%)\n' % (gem, gemdec, 100 * (gem + 39 * gemdec) / whole))
open(dst, 'w').write(''.join(txt))
print(''.join(txt))
