"""Per-kernel opcode histogram (executed warp instructions) from `ncu --page source --csv --print-source sass`."""
import csv, sys, collections
fn = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern = None; hdr = None; hist = None
def flush():
    if kern and hist:
        tot = sum(hist.values())
        print("== %s   total %.1fM warp-instr" % (kern[:90], tot / 1e6))
        for op, n in hist.most_common(top):
            print("   %-14s %8.1fM  %5.1f%%" % (op, n / 1e6, 100.0 * n / tot))
for row in csv.reader(open(fn)):
    if not row: continue
    if row[0] == "Kernel Name":
        flush(); kern = row[1]; hdr = None; hist = collections.Counter(); continue
    if row[0] == "Address":
        hdr = {k: i for i, k in enumerate(row)}; continue
    if hdr is None: continue
    try:
        n = float(row[hdr["Instructions Executed"]])
    except Exception:
        continue
    src = row[hdr["Source"]].strip()
    toks = src.split()
    if toks and toks[0].startswith("@"): toks = toks[1:]
    op = toks[0].split(".")[0] if toks else "?"
    full = toks[0] if toks else "?"
    key = op if op not in ("LDS", "STS", "LDG", "STG") else full
    hist[key] += n
flush()
