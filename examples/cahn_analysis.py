#!/usr/bin/env python3
"""Coarsening statistics of a Cahn-Hilliard run from its snapshots: the analysis of the reference's
cuPentCahnADI/plotting.py:18-90 on the raw-binary snapshots examples/cuPentCahnADI writes (custen_cahn_write_snapshot).

For every snapshot c(t):
    s(t)   = 1 / (1 - <c^2>),   <c^2> = (1 / (2 pi)^2) * Simpson integral of c^2 over [0, 2 pi]^2        (plotting.py:62-66)
    1/k1   = sum(|k|^-1 |c_hat(k)|^2) / sum(|c_hat(k)|^2) over the modes with kx != 0 and ky != 0        (plotting.py:41-56, 68-76)
Both follow t^(1/3) in the coarsening regime.  Prints a table (and writes it as CSV with --csv); draws the reference's
log-log figure when matplotlib is available and --plot is given.

    python examples/cahn_analysis.py output/ [--csv analysis.csv] [--plot analysis.png]
"""
import argparse
import math
import os
import re
import struct

import numpy as np

MAGIC = b"CUSTENC1"


def write_snapshot(directory, time, field):
    """numpy twin of custen_cahn_write_snapshot (custen_b200/csrc/cahn.cu): same name, same bytes."""
    field = np.ascontiguousarray(field, dtype="<f8")
    path = os.path.join(directory, "cahn_hilliard_%0.10f.bin" % time)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<qqd", field.shape[1], field.shape[0], time))
        f.write(field.tobytes())
    return path


def read_snapshot(path):
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a cuSten-B200 Cahn-Hilliard snapshot")
        nx, ny, time = struct.unpack("<qqd", f.read(24))
        data = np.frombuffer(f.read(nx * ny * 8), dtype="<f8")
    if data.size != nx * ny:
        raise ValueError(f"{path}: truncated")
    return time, data.reshape(ny, nx)


def simpson(y, x):
    """Composite Simpson rule along the last axis for equally spaced x; an odd number of intervals is closed with the
    trapezoid rule on the last one.  (plotting.py uses scipy.integrate.simps on np.linspace samples.)"""
    n = y.shape[-1]
    h = (x[-1] - x[0]) / (n - 1)
    m = n if n % 2 == 1 else n - 1
    w = np.ones(m)
    w[1:-1:2] = 4.0
    w[2:-1:2] = 2.0
    total = (h / 3.0) * np.tensordot(y[..., :m], w, axes=([-1], [0]))
    if m != n:
        total = total + 0.5 * h * (y[..., -2] + y[..., -1])
    return total


def statistics(c):
    """(s, 1 / k1) of one field, the reference's expressions."""
    ny, nx = c.shape
    x = np.linspace(0.0, 2.0 * math.pi, nx)
    y = np.linspace(0.0, 2.0 * math.pi, ny)
    avg = (1.0 / ((2.0 * math.pi) ** 2)) * simpson(simpson(np.square(c), y), x)
    s = 1.0 / (1.0 - avg)
    k = np.fft.fftfreq(nx, 1.0 / nx)
    kx, ky = np.meshgrid(k[1:], np.fft.fftfreq(ny, 1.0 / ny)[1:])
    mod_k_inv = 1.0 / np.sqrt(np.square(kx) + np.square(ky))
    ft = np.square(np.abs(np.fft.fft2(c)))[1:ny, 1:nx]
    return s, float(np.sum(mod_k_inv * ft) / np.sum(ft))


def analyse(directory):
    """Rows (t, s(t), 1/k1, t^(1/3)) for every snapshot in `directory`, ordered by time."""
    rows = []
    for name in os.listdir(directory):
        if not re.fullmatch(r"cahn_hilliard_\d+\.\d+\.bin", name):
            continue
        t, c = read_snapshot(os.path.join(directory, name))
        s, k1 = statistics(c)
        rows.append((t, s, k1, t ** (1.0 / 3.0)))
    rows.sort()
    return rows


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("directory")
    ap.add_argument("--csv")
    ap.add_argument("--plot")
    args = ap.parse_args()
    rows = analyse(args.directory)
    print(f"{'t':>14} {'s(t)':>14} {'1/k1':>14} {'t^(1/3)':>14}")
    for r in rows:
        print(" ".join(f"{v:14.8f}" for v in r))
    if args.csv:
        with open(args.csv, "w") as f:
            f.write("t,s,inv_k1,t_third\n")
            for r in rows:
                f.write(",".join(repr(float(v)) for v in r) + "\n")
    if args.plot:
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except ImportError:
            print("matplotlib is not installed: no figure")
            return
        t = [r[0] for r in rows]
        plt.loglog(t, [r[1] for r in rows], label="s(t)")
        plt.loglog(t, [r[3] for r in rows], label="t^{1/3}")
        plt.loglog(t, [r[2] for r in rows], label="1 / k_1")
        plt.legend(loc="upper left")
        plt.xlabel("t")
        plt.savefig(args.plot, dpi=300)


if __name__ == "__main__":
    main()
