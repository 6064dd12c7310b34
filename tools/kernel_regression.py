"""Per-kernel regression guard over tools/kernel_bench.py logs (runs on the CPU box).

    python tools/kernel_regression.py <new_kernel_bench.log> [<baseline.log>] [--tol 0.05]

Every kernel line of a log ("name   123.4 us ...") is normalised by the BOX FACTOR of its log: the median, over the torch
controls measured in the same run (the indented "(torch ...)" lines: library GEMMs / layer_norm / SDPA on the same box
minutes apart), of control_new / control_baseline. That removes most of the box-to-box and thermal spread (+-5 % on this
pool) that made round 1's logs look like a regression of the attention forward. (A single control line is itself only good
to +-10 %, so the kernel's own control is printed for information and the median decides.) A kernel whose normalised time
grew by more than --tol (default 5 %) against the baseline is reported and the exit code is 1.
Without a baseline argument the last kept profiles/*kernel_bench*.log (by name) that is not the new file is used."""
import re
import sys
from pathlib import Path

LINE = re.compile(r"^(\s*)(\(?[^\d].*?\)?)\s+([\d.]+) us\b")


def parse(path):
    """-> (kernels: {name: us}, control_of: {name: control name or None}, controls: {name: us})"""
    kernels, control_of, controls = {}, {}, {}
    pending = []
    ctl_count = {}
    for raw in Path(path).read_text().splitlines():
        m = LINE.match(raw)
        if not m:
            continue
        indent, name, us = m.group(1), m.group(2).strip(), float(m.group(3))
        if name.startswith("("):
            ctl_count[name] = ctl_count.get(name, 0) + 1
            key = f"{name}#{ctl_count[name]}"
            controls[key] = us
            for k in pending:
                control_of[k] = key
            pending = []
        else:
            if name in kernels:
                continue
            kernels[name] = us
            control_of[name] = None
            pending.append(name)
            if not indent and re.match(r"^(dgrad|layernorm bwd|attention bwd|colsum|im2col|cast|assemble|optimizer)", name):
                pending.remove(name)  # these have no control of their own
    return kernels, control_of, controls


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tol = 0.05
    if "--tol" in sys.argv:
        tol = float(sys.argv[sys.argv.index("--tol") + 1])
        args = [a for a in args if a != str(tol) and a != sys.argv[sys.argv.index("--tol") + 1]]
    if not args:
        print(__doc__)
        return 2
    new = Path(args[0])
    if len(args) > 1:
        base = Path(args[1])
    else:
        kept = sorted(p for p in (Path(__file__).resolve().parents[1] / "profiles").glob("*kernel_bench*.log") if p.resolve() != new.resolve() and "modes" not in p.name and "gemm" not in p.name)
        if not kept:
            print("no baseline log under profiles/")
            return 2
        base = kept[-1]
    nk, nctl, nc = parse(new)
    bk, bctl, bc = parse(base)
    common_ctl = [c for c in nc if c in bc]
    ratios = sorted(nc[c] / bc[c] for c in common_ctl)
    box = (ratios[(len(ratios) - 1) // 2] + ratios[len(ratios) // 2]) / 2 if ratios else 1.0
    print(f"new {new.name} vs baseline {base.name}; box factor (median of {len(common_ctl)} torch controls): {box:.3f}")
    bad = []
    for name, us in nk.items():
        if name not in bk:
            continue
        c = nctl.get(name)
        own = f"x{nc[c] / bc[c]:.3f}" if c is not None and c in bc and bctl.get(name) == c else "   -  "
        ratio = (us / bk[name]) / box
        flag = "  <-- REGRESSION" if ratio > 1 + tol else ""
        print(f"{name:36s} {bk[name]:8.1f} -> {us:8.1f} us   own control {own}   normalised x{ratio:.3f}{flag}")
        if flag:
            bad.append(name)
    if bad:
        print(f"FAIL: {len(bad)} kernel(s) more than {tol:.0%} slower than the baseline relative to the torch controls: {bad}")
        return 1
    print("OK: no kernel regressed")
    return 0


if __name__ == "__main__":
    sys.exit(main())
