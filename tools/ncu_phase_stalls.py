"""Aggregate an `ncu --page source --csv` export of the fused kernel into per-phase rows (phases are delimited by the
CTA barriers in SASS order): samples, warp instructions, shared-memory wavefronts and the top stall reasons.
    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_phase_stalls.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
new = lambda: {'n': 0, 'samples': 0, 'inst': 0, 'st': {s: 0 for s in stalls}, 'sh_excess': 0, 'sh_wave': 0, 'local': 0}
regions = []; cur = new()
for r in data:
    src = r[idx['Source']]
    smp = int(r[idx['# Samples']] or 0)
    cur['n'] += 1; cur['samples'] += smp; cur['inst'] += int(r[idx['Instructions Executed']] or 0)
    cur['sh_excess'] += int(r[idx['L1 Wavefronts Shared Excessive']] or 0); cur['sh_wave'] += int(r[idx['L1 Wavefronts Shared']] or 0)
    if 'LDL' in src or 'STL' in src: cur['local'] += int(r[idx['Instructions Executed']] or 0)
    for s in stalls: cur['st'][s] += int(r[idx[s]] or 0)
    if 'BAR.SYNC' in src:
        regions.append(cur); cur = new()
regions.append(cur)
tot = sum(r['samples'] for r in regions)
names = (sys.argv[2].split(',') if len(sys.argv) > 2 else ['init', 'stage', 'sincos', 'chains', 'feet', 'sblocks', 'chol', 'wcols', 'proj', 'fill', 'M', 'tail'])
for i, r in enumerate(regions):
    top = sorted(r['st'].items(), key=lambda kv: -kv[1])[:5]
    print('%-8s sass=%5d samples=%7d (%4.1f%%) warp-inst=%9d local=%7d shwave=%8d excess=%8d  %s' % (
        names[i] if i < len(names) else str(i), r['n'], r['samples'], 100 * r['samples'] / max(tot, 1), r['inst'], r['local'], r['sh_wave'], r['sh_excess'],
        ', '.join('%s=%d' % (k.replace('stall_', ''), v) for k, v in top)))
