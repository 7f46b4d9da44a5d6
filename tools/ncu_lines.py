"""Aggregates an ncu report's warp-stall samples per CUDA source line (needs -lineinfo + --import-source).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [launch_index] [top]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# split into kernel launches at "Function Name" markers
launch, cur_file, hdr, launches = -1, None, None, collections.defaultdict(lambda: collections.defaultdict(lambda: [0, collections.Counter(), ""]))
seen_first_file = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        stalls = [(j, c[6:]) for j, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        if seen_first_file is None:
            seen_first_file = cur_file
        if cur_file == seen_first_file:
            launch += 1
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        s = int(r[si])
    except ValueError:
        continue
    if r[0] != "":
        cur_line, cur_src = int(r[0]), r[1].strip()
    key = (cur_file, cur_line)
    e = launches[launch][key]
    e[0] += s
    e[2] = cur_src
    for j, nme in stalls:
        v = int(r[j] or 0)
        if v:
            e[1][nme] += v
d = launches[which]
tot = sum(e[0] for e in d.values()) or 1
print("launch %d of %d matching %s: %d samples" % (which, len(launches), kern, tot))
for (f, ln), e in sorted(d.items(), key=lambda x: -x[1][0])[:top]:
    st = " ".join("%s=%d" % (k, v) for k, v in e[1].most_common(3))
    print("%5.1f%% %s:%-4d %-84s %s" % (100.0 * e[0] / tot, f[:12], ln, e[2][:84], st))
