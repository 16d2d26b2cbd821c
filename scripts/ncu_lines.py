#!/usr/bin/env python3
"""Per-source-line share of executed instructions / stall samples for one kernel of an .ncu-rep."""
import csv
import subprocess
import sys


def main(path, kernel, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass,cuda',
                          '--kernel-name', f'regex:{kernel}'], capture_output=True, text=True).stdout
    cur, data = None, []
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) > 8 and r[0] not in ('', 'Line No'):
            try:
                data.append((int(r[7]), int(r[6]), cur, r[0], r[1][:110]))
            except ValueError:
                pass
    tot = sum(d[0] for d in data) or 1
    tots = sum(d[1] for d in data) or 1
    print('total warp instructions', tot, 'samples', tots)
    for d in sorted(data, reverse=True)[:top]:
        print(f'{d[0] / tot * 100:5.1f}% inst {d[1] / tots * 100:5.1f}% smp  {d[2]}:{d[3]:>4s} {d[4]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
