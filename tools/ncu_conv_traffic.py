"""Sum the DRAM traffic of the conv_tc_kernel launches of ONE bench step from an ncu csv log
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:conv_tc --csv --log-file gpurun_out/conv_dram.csv python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline
    python tools/ncu_conv_traffic.py gpurun_out/conv_dram.csv 83 8 > profiles/conv_dram_r01.json
(83 = conv launches per step, 8 = clips per step; the LAST 83 launches of the log are the roofline pass.)"""
import csv
import json
import sys


def main():
    path, per_step, clips = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith('==')))
    by_id = {}
    for r in rows:
        d = by_id.setdefault(int(r['ID']), {})
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1}.get(u, 1)
        d[r['Metric Name']] = v * scale
    ids = sorted(by_id)[-per_step:]
    rd = sum(by_id[i].get('dram__bytes_read.sum', 0) for i in ids)
    wr = sum(by_id[i].get('dram__bytes_write.sum', 0) for i in ids)
    t = sum(by_id[i].get('gpu__time_duration.sum', 0) for i in ids)
    print(json.dumps(dict(kernel='conv_tc_kernel', launches=len(ids), clips_per_step=clips, dram_bytes_read=rd, dram_bytes_write=wr,
                          dram_bytes=rd + wr, gpu_time_s_under_ncu=t,
                          source='ncu dram__bytes_read.sum + dram__bytes_write.sum over the conv_tc launches of one bench step (%s)' % path),
                     indent=1))


if __name__ == '__main__':
    main()
