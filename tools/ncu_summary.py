"""Summarise an `ncu --set full` report (.ncu-rep) into a text file for profiles/ and update profiles/roofline_traffic.json.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x_ncu_full_summary.txt ["note"]

Per captured launch: duration, DRAM bytes (read / write / sum), achieved DRAM GB/s, tensor-pipe and SM throughput,
registers, shared memory.  bench.py's `roofline.traffic` reads the json this script maintains (kernel name -> mean
dram__bytes_read.sum + dram__bytes_write.sum per launch).
"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'launch__grid_size']

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
        'ns': 1e-9, 'nsecond': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'ms': 1e-3, 'msecond': 1e-3, 's': 1.0, 'second': 1.0}


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(hdr):
        for w in WANT:
            if h.endswith(w) and w not in col:
                col[w] = i
    kcol = hdr.index('Kernel Name')
    out = [note, f'# source: {os.path.basename(rep)} (ncu --set full --clock-control none); per captured launch']
    traffic = {}
    for r in data:
        if len(r) <= kcol:
            continue
        name = r[kcol]
        vals = {}
        for w, i in col.items():
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            vals[w] = v * UNIT.get(units[i], 1.0)
        dur = vals.get('gpu__time_duration.sum', 0.0)
        rd, wr = vals.get('dram__bytes_read.sum', 0.0), vals.get('dram__bytes_write.sum', 0.0)
        out.append(f'kernel: {name[:110]}')
        out.append(f'  duration {dur * 1e6:.1f} us | DRAM read {rd / 1e6:.2f} MB + write {wr / 1e6:.2f} MB = {(rd + wr) / 1e6:.2f} MB'
                   f' -> {((rd + wr) / dur / 1e9) if dur else 0:.0f} GB/s')
        for w in WANT[3:]:
            if w in vals:
                out.append(f'  {w} = {vals[w]:.6g}')
        traffic.setdefault(name.split('(')[0].strip(), []).append(rd + wr)
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))
    jpath = os.path.join(os.path.dirname(os.path.abspath(dst)), 'roofline_traffic.json')
    recs = []
    if os.path.exists(jpath):
        recs = json.load(open(jpath))
    for k, v in traffic.items():
        recs = [x for x in recs if x.get('kernel') != k]
        recs.append({'kernel': k, 'dram_bytes_per_launch': sum(v) / len(v), 'launches': len(v),
                     'source': os.path.basename(dst)})
    json.dump(recs, open(jpath, 'w'), indent=1)


main()
