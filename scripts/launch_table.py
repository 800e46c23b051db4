"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table of kernels by total time."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows:
    if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("adpst::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * (1e3 if r[ui] in ("ms", "msecond") else 1.0)
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
total = sum(tot.values())
print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (k[:60], cnt[k], v, 100 * v / total))
print("| **total** | %d | %.1f | |" % (sum(cnt.values()), total))
