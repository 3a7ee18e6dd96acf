"""Executed warp instructions of mid2_kernel per phase of the kernel body, from one ncu --set full --import-source capture.

  cuobjdump -xelf all hdgnn_b200/build/k_mid.o && nvdisasm -gi -c k_mid.sm_100a.cubin > sass_lines.txt
  ncu -i prof.ncu-rep --page source --csv --kernel-name regex:mid2 > sass_metrics.csv
  python tools/phase_attrib.py sass_lines.txt sass_metrics.csv '<mangled kernel name>' out.csv

Every SASS instruction is attributed to the OUTERMOST mid2.cuh line of its inline chain (the call site in the kernel body), and
the lines are grouped by the M2_PHASE(i) marker that precedes them in the source."""
import bisect, collections, csv, os, re, sys

lines_txt, metrics_csv, mangled, out = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {0: "A prologue: inputs, hunk sort, attribute sort, transposed bitmap, classes (before the wait)", 1: "weights + derived tables",
         16: "entity pair layer forward (class tables / search + edge walk)", 17: "B entity-state MLP forward", 2: "D pooling forward",
         13: "D/K pooling (ident)", 14: "D pooling (general tail)", 15: "D segmented reduce by hunk", 3: "hunk tables + E hunk pair sweep forward",
         4: "F second layer + head tables", 5: "G1 relation head: logits, softmax, CE", 6: "G2 delta sums sweep", 7: "H head backward (node level)",
         8: "I hunk pair sweep backward", 9: "J hunk first-layer gradients, d/dnb", 12: "K pooling backward", 10: "L entity-state MLP backward",
         18: "entity pair layer backward: tables / scans", 19: "entity pair layer backward: items + reduction", 11: "end", 20: "general: edge walks",
         21: "general: edge walks"}
loc, off2line, inside = None, {}, False
for l in open(lines_txt, errors="ignore"):
    if l.startswith(".text."):
        inside = mangled in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        outer = [ln for f, ln in chain if f.endswith("mid2.cuh") and ln >= 250]
        loc = outer[-1] if outer else None
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m:
        off2line[int(m.group(1), 16)] = loc
src = open(os.path.join(ROOT, "hdgnn_b200", "csrc", "mid2.cuh")).read().split("\n")
marks = [(i + 1, int(re.search(r"M2_PHASE\((\d+)", s).group(1))) for i, s in enumerate(src) if "M2_PHASE(" in s and "define" not in s]
bounds = [m[0] for m in marks]
rows = [r for r in csv.reader(open(metrics_csv, errors="ignore")) if r]
h0 = next(i for i, r in enumerate(rows) if r[0] == "Address")
h = rows[h0]
ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
data = []
for r in rows[h0 + 1:]:
    if r[0] == "Address" or r[0] == "Kernel Name":
        break
    data.append(r)
base = int(data[0][ia], 16)
inst, samp = collections.Counter(), collections.Counter()
for r in data:
    try:
        a, n, s = int(r[ia], 16) - base, int(r[ie]), int(r[isamp])
    except (ValueError, IndexError):
        continue
    ln = off2line.get(a)
    if ln is None:
        key = "(helpers / unattributed)"
    else:
        k = bisect.bisect_right(bounds, ln)
        key = NAMES.get(marks[k - 1][1], f"phase {marks[k - 1][1]}") if k > 0 else "kernel entry, layout, weight-load lambda"
    inst[key] += n; samp[key] += s
tot, tots = sum(inst.values()), max(1, sum(samp.values()))
with open(out, "w") as f:
    f.write("phase,warp_instructions_executed,share,stall_samples_share\n")
    for k, v in sorted(inst.items(), key=lambda kv: -kv[1]):
        f.write(f"\"{k}\",{v},{v / tot:.3f},{samp[k] / tots:.3f}\n")
    f.write(f"\"TOTAL (one launch)\",{tot},1.000,1.000\n")
print(open(out).read())
