"""Per-instruction stall samples of one .ncu-rep capture (source page), aggregated by SASS region.
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows[:40]) if len(r) > 5 and 'Source' in r][0]
    h = rows[hi]
    si, sa, ex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    body = []
    for k, r in enumerate(rows[hi + 1:]):
        try:
            body.append((k, int(r[sa] or 0), int(r[ex] or 0), r[si]))
        except (ValueError, IndexError):
            pass
    tot = sum(b[1] for b in body)
    print('total samples', tot, 'instructions', len(body))
    for b in sorted(body, key=lambda b: -b[1])[:topn]:
        print('%5d %5d %8d  %s' % b)
    return body


if __name__ == '__main__':
    main()
