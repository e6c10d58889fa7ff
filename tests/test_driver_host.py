"""CPU: host-side pieces of main.jl that surround the hot path (hyperelasticsolver_b200/driver.py)."""
import random

import numpy as np

from hyperelasticsolver_b200.driver import HEADER, get_filename, julia_float, read_data


def test_julia_float_format():
    cases = {0.1: "0.1", 9.066365117490715: "9.066365117490715", 1e-5: "1.0e-5", 1.5e-7: "1.5e-7", 0.0001: "0.0001",
             0.00012: "0.00012", 100000.0: "100000.0", 1e6: "1.0e6", 1234567.0: "1.234567e6", -2.5: "-2.5", 300.0: "300.0",
             1e22: "1.0e22", 123456.789: "123456.789", 5e-324: "5.0e-324", 1 / 3: "0.3333333333333333", 2.0: "2.0",
             10.0: "10.0", 0.5: "0.5", 9.999999e5: "999999.9", 0.001004474742147317: "0.001004474742147317"}
    for x, s in cases.items():
        assert julia_float(x) == s, (x, julia_float(x), s)
    assert julia_float(0.0) == "0.0" and julia_float(-0.0) == "-0.0"
    assert julia_float(float("nan")) == "NaN" and julia_float(float("inf")) == "Inf" and julia_float(float("-inf")) == "-Inf"
    rnd = random.Random(1)
    for _ in range(20000):
        x = rnd.uniform(-1, 1) * 10 ** rnd.randint(-12, 12)
        assert float(julia_float(x)) == x


def test_csv_names_and_read(tmp_path):
    assert get_filename(0) == "sol_000000.csv" and get_filename(641) == "sol_000641.csv"   # main.jl:108
    cols = HEADER.split("\t")
    assert len(cols) == 30 and cols[0] == "a1" and cols[15] == "a2" and cols[5] == "S1" and cols[6] == "F111"
    P = np.random.default_rng(0).normal(size=(7, 30))
    f = tmp_path / get_filename(3)
    with open(f, "w") as io:
        io.write(HEADER + "\n")
        for row in P:
            io.write("\t".join(julia_float(float(v)) for v in row) + "\n")
    P2, nx = read_data(str(f))
    assert nx == 7 and np.array_equal(P, P2)          # shortest round-trip text is lossless
    # plotter.py reads these columns: 0/15 alpha, 1/16 rho, 2-4/17-19 u, 5/20 S
    data = np.loadtxt(str(f), skiprows=1)
    assert np.array_equal(data[:, [0, 15]], P[:, [0, 15]])
