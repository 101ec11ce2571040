"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (us, launches, share)."""
import collections, csv, re, sys

def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1; agg[name][1] += v; tot += v
    return agg, tot

if __name__ == "__main__":
    agg, tot = load(sys.argv[1])
    print("| kernel | launches | total us | us/launch | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:80]}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |")
    print(f"\ntotal {tot:.0f} us over {sum(n for n, _ in agg.values())} launches")
