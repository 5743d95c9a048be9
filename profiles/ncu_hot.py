#!/usr/bin/env python
"""Hottest SASS instructions of each kernel in an .ncu-rep (source page): executed count and stall samples.
python profiles/ncu_hot.py <rep> [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; hdr = rows[i+1]; i += 2
        body = []
        while i < len(rows) and not (rows[i] and rows[i][0] == 'Kernel Name'):
            if len(rows[i]) == len(hdr): body.append(dict(zip(hdr, rows[i])))
            i += 1
        tot_inst = sum(int(b['Instructions Executed']) for b in body)
        tot_samp = sum(int(b['# Samples']) for b in body)
        print('==', name, 'SASS lines', len(body), 'warp-insts', tot_inst, 'samples', tot_samp)
        # opcode histogram weighted by executed count
        ops = {}
        for b in body:
            op = b['Source'].split()[0] if not b['Source'].strip().startswith('@') else b['Source'].split()[1]
            op = op.split('.')[0]
            ops[op] = ops.get(op, 0) + int(b['Instructions Executed'])
        print('   opcode mix:', ', '.join('%s %.1f%%' % (k, 100.0*v/tot_inst) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
        for b in sorted(body, key=lambda b: -int(b['# Samples']))[:top]:
            stalls = {k: int(v) for k, v in b.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v) > 0}
            main = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
            print('   %6.2f%% samp %5.2f%% inst  %-60s %s' % (100.0*int(b['# Samples'])/max(tot_samp,1), 100.0*int(b['Instructions Executed'])/tot_inst, b['Source'].strip()[:60], main))
    else:
        i += 1
