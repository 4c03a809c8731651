"""Summarise an `ncu --set full` capture of the tensor-core conv launches (raw page exported with
`ncu -i X.ncu-rep --page raw --csv`): one line per launch (duration, DRAM bytes, tensor-pipe activity) plus totals,
and the small JSON bench.py reads for `roofline.traffic` / `tensor_pipe_active_pct_ncu`.

usage: python tools/ncu_summarise.py RAW.csv OUT.txt [OUT.json] [--title "..."] [--shape "[16,624,1024,3]"]"""
import csv
import json
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'second': 1e6, '%': 1.0}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    title = sys.argv[sys.argv.index('--title') + 1] if '--title' in sys.argv else ''
    if '--title' in sys.argv:
        args.remove(title)
    shape = sys.argv[sys.argv.index('--shape') + 1] if '--shape' in sys.argv else '[8,624,1024,3]'
    if '--shape' in sys.argv:
        args.remove(shape)
    raw, out_txt = args[0], args[1]
    out_json = args[2] if len(args) > 2 else None
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = col[name]
        return float(r[i].replace(',', '')) * UNIT[units[i]]

    lines, tot_us, tot_rd, tot_wr, tw = [], 0.0, 0.0, 0.0, 0.0
    for r in data:
        name = r[col['Kernel Name']].replace('void ', '')
        us = val(r, 'gpu__time_duration.sum')
        rd, wr = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
        tp = val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
        lines.append('%-44s %8.1f us  dram rd %7.1f MB wr %7.1f MB (%.2f TB/s)  tensor pipe active %5.1f %%'
                     % (name[:44], us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e6, tp))
        tot_us += us
        tot_rd += rd
        tot_wr += wr
        tw += tp * us
    n = len(data)
    with open(out_txt, 'w') as f:
        if title:
            f.write(title + '\n')
        f.write('\n'.join(lines) + '\n')
        f.write('%d launches: %.3f ms, dram read %.1f MB + write %.1f MB = %.1f MB per launch; time-weighted tensor pipe '
                'active %.1f %%\n' % (n, tot_us / 1e3, tot_rd / 1e6, tot_wr / 1e6, (tot_rd + tot_wr) / n / 1e6, tw / tot_us))
    if out_json:
        json.dump({'conv_tc_dram_bytes_per_launch': (tot_rd + tot_wr) / n, 'launches': n,
                   'workload': 'dense %s network pass (tools/prof_forward.py), %d tensor-core conv launches' % (shape, n),
                   'dram_read_bytes_total': tot_rd, 'dram_write_bytes_total': tot_wr, 'ncu_time_ms_total': tot_us / 1e3,
                   'conv_tc_tensor_pipe_active_pct_time_weighted': tw / tot_us,
                   'source': out_txt + ' (ncu --set full, --clock-control none)'}, open(out_json, 'w'), indent=1)
    print(open(out_txt).read().splitlines()[-1])


if __name__ == '__main__':
    main()
