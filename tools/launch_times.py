"""print (kernel, ms) of an ncu launch list (--metrics gpu__time_duration.sum --csv); optional substring filter"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
flt = sys.argv[2:] 
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i
        break
else:
    print(open(sys.argv[1]).read()[-800:]); sys.exit(1)
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[start + 1:]:
    if len(r) <= vi: continue
    name = r[ki].replace('plaidgpu::', '').replace('<unnamed>::', '').replace('void ', '')[:44]
    if flt and not any(f in name for f in flt): continue
    try: print(f"{name:46s} {float(r[vi].replace(',', '')) / 1e6:9.3f} ms")
    except ValueError: pass
