"""per-source-line hot spots of one kernel from an .ncu-rep (needs --import-source on, -lineinfo):
python tools/ncu_hotlines.py rep.ncu-rep kernel_regex [top_n] [smp|inst]"""
import csv, subprocess, sys, io, os
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
key = 1 if (len(sys.argv) > 4 and sys.argv[4] == "inst") else 0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern, '--print-source',
                      'cuda,sass'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True).stdout
fname, hdr, lines = "?", None, []
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
    elif r[0] == "Line No":
        hdr = r
        i_smp, i_ins = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and r[0].isdigit():
        try:
            lines.append((float(r[i_smp] or 0), float(r[i_ins] or 0), "%s:%s" % (fname, r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
ts = sum(l[0] for l in lines) or 1.0
ti = sum(l[1] for l in lines) or 1.0
print("kernel %s: %.4g warp instructions, %.4g stall samples" % (kern, ti, ts))
for smp, ins, loc, src in sorted(lines, key=lambda l: -l[key])[:top]:
    print("%5.1f%% smp %5.1f%% inst | %-22s | %s" % (100 * smp / ts, 100 * ins / ti, loc, src))
