"""Joins an `ncu --page source --csv` export (per-SASS-instruction executed counts and stall samples) with the line table
of the cubin (`nvdisasm -g`), and prints the executed warp-instructions per launch and the stall samples per SOURCE LINE
of one kernel -- where the instructions of the tree kernel go, in terms of mcts.cu.

    cuobjdump -xelf all betazero_b200/libbetazero_b200.so          # -> mcts.sm_100a.cubin
    nvdisasm -g mcts.sm_100a.cubin > mcts_g.txt
    ncu -i X.ncu-rep --page source --csv > src.csv
    python profiles/sass_by_line.py src.csv mcts_g.txt step_wave_kernelILi0ELi8 [warps]
"""
import csv, re, sys, collections

src_csv, disasm, kernel = sys.argv[1:4]
warps = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)
ex = {int(r[ix["Address"]], 16) - base: (int(r[ix["Instructions Executed"]]), int(r[ix["Warp Stall Sampling (All Samples)"]])) for r in data}

line_of, cur, on = {}, None, False
for ln in open(disasm):
    if ln.startswith(".text.") and kernel in ln:
        on = True
        continue
    if on and ln.startswith("//---------------------"):
        break
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur

agg = collections.defaultdict(lambda: [0, 0, 0])
for off, (e, s) in ex.items():
    a = agg[line_of.get(off)]
    a[0] += e
    a[1] += s
    a[2] += 1
tot_e = sum(a[0] for a in agg.values())
tot_s = sum(a[1] for a in agg.values())
print(f"# {kernel}: {tot_e / warps:.0f} warp-instructions per warp, {tot_s} stall samples")
srcs = {}
def text(key):
    if key is None:
        return "?"
    f, l = key
    if f not in srcs:
        try:
            srcs[f] = open(f"betazero_b200/csrc/{f}").read().split("\n")
        except OSError:
            srcs[f] = []
    return srcs[f][l - 1].strip()[:90] if l - 1 < len(srcs[f]) else ""
for key, a in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if a[0] == 0 and a[1] == 0:
        continue
    name = f"{key[0]}:{key[1]}" if key else "?"
    print(f"{name:22s} exec/warp {a[0] / warps:7.1f} ({100 * a[0] / tot_e:4.1f}%)  stalls {100 * a[1] / max(tot_s, 1):4.1f}%  | {text(key)}")
