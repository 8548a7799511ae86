"""Join an ncu SASS source page (`ncu -i X.ncu-rep --page source --csv`) with `nvdisasm
--print-line-info` of the same cubin and aggregate executed instructions / stall samples per
CUDA source line.  usage: ncu_by_line.py src.csv all.sass <mangled kernel name> [top]"""
import collections
import csv
import re
import sys


def load_lines(sass_path, kernel):
    lines = open(sass_path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("//---") and ".text." + kernel + " " in l)
    cur = ("?", 0)
    out = {}
    for l in lines[start + 1:]:
        if l.startswith("//---"):
            break
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    src, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    linemap = load_lines(sass, kernel)
    rows = list(csv.reader(open(src)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ia, ii, is_ = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    by_line = collections.defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= ii or not r[ia].startswith("0x"):
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        key, _ = linemap.get(addr - base, (("?", 0), ""))
        n, s = int(r[ii] or 0), int(r[is_] or 0)
        by_line[key][0] += n
        by_line[key][1] += s
        tot_i += n
        tot_s += s
    print("total warp-instructions %d, stall samples %d" % (tot_i, tot_s))
    srcs = {}
    for (f, ln), (n, s) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            try:
                srcs[f] = open("/root/repo/gym_comm_b200/csrc/" + f).read().split("\n")
            except Exception:
                srcs[f] = []
        text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
        print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100.0 * n / tot_i, 100.0 * s / max(tot_s, 1), f, ln, text))


if __name__ == "__main__":
    main()
