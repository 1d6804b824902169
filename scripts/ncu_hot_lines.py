"""Top source lines of a kernel by warp-stall samples, from an .ncu-rep captured with
--import-source on (and a -lineinfo build):  python scripts/ncu_hot_lines.py rep [N]"""
import csv, subprocess, sys, collections

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    acc = collections.OrderedDict()
    fpath, hdr = "", None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]; hdr = None; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ci = r.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr is None or len(r) <= ci:
            continue
        try:
            n = float(r[ci].replace(",", "") or 0)
        except ValueError:
            continue
        if r[0]:
            cur = (fpath, r[0], r[1].strip()[:100])
        acc[cur] = acc.get(cur, 0) + n
    tot = sum(acc.values()) or 1
    print(f"# {rep}: {int(tot)} warp-stall samples; top {top} source lines")
    for (f, ln, src), n in sorted(acc.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{100 * n / tot:5.1f}%  {f}:{ln:<5} {src}")

if __name__ == "__main__":
    main()
