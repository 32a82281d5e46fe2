"""Hot CUDA source lines of one kernel from an Nsight Compute report captured with --import-source on (run here, no GPU):
    python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    fpath, hdr, out, seen_fn = None, None, [], None
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == 'File Path':
            fpath = r[1]
            continue
        if r[0] == 'Function Name':
            if seen_fn is None:
                seen_fn = r[1]
            elif r[1] != seen_fn and fpath is None:
                break
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr and r[0].isdigit():
            si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
            try:
                out.append((fpath.split('/')[-1], int(r[0]), r[1].strip(), int(r[si] or 0), int(r[ii] or 0)))
            except (ValueError, IndexError):      # source text with embedded quotes breaks ncu's CSV row: take the counters from the tail
                k = len(hdr) - si
                try:
                    out.append((fpath.split('/')[-1], int(r[0]), r[1].strip(), int(r[len(r) - k] or 0), int(r[len(r) - k + 1] or 0)))
                except (ValueError, IndexError):
                    pass
    ts, ti = sum(o[3] for o in out) or 1, sum(o[4] for o in out) or 1
    print(f'kernel {seen_fn}: {ts} stall samples, {ti} warp instructions')
    for f, ln, src, s, i in sorted(out, key=lambda o: -o[3])[:top]:
        print(f'{f:20s} L{ln:<5d} samples {100 * s / ts:5.1f}%  inst {100 * i / ti:5.1f}%  {src[:120]}')


if __name__ == '__main__':
    main()
