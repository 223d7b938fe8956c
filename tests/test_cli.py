"""The C++ deck reader / writer (cli/): every shipped deck parses to exactly
what the oracle-side restatement of read_input yields (bit for bit), the Fortran edit
descriptors are reproduced, the headers have the reference's layout; on a GPU the whole
command line produces the reference's output file from the C-ABI results."""
import os
import subprocess

import numpy as np
import pytest

import unconfined_b200 as ub
from oracle import deck
from helpers import ROOT

CFG = os.path.join(ROOT, "configs")
CLI = os.path.join(ROOT, "cli", "unconfined_cli")
DECKS = sorted(f for f in os.listdir(CFG) if f.endswith("-input.dat") or f.endswith(".in"))


@pytest.fixture(scope="module")
def cli():
    ub.lib()                       # the CUDA library must exist (built by __graft_entry__.build())
    subprocess.run(["make", "-C", os.path.join(ROOT, "cli")], check=True, capture_output=True)
    return CLI


def run(cli, *args, cwd=CFG):
    r = subprocess.run([cli, *args], capture_output=True, text=True, cwd=cwd)
    assert r.returncode == 0, r.stderr
    return r.stdout


def parse_dump(text):
    out = {}
    for line in text.strip().split("\n"):
        k, *v = line.split()
        out[k] = v
    return out


def hexv(tokens):
    return np.array([float.fromhex(t) for t in tokens[1:]])


@pytest.mark.parametrize("name", DECKS)
def test_deck_reader_matches_oracle_reader_bitwise(cli, name):
    d = deck.read_deck(os.path.join(CFG, name))
    g = parse_dump(run(cli, name, "--dump"))
    assert int(g["model"][0]) == d["model"] and int(g["M"][0]) == d["M"]
    assert [int(x) for x in g["j0s"]] == list(d["j0s"])
    assert int(g["nacc"][0]) == d["gl_nacc"] and int(g["ord"][0]) == d["gl_ord"]
    assert int(g["ts_k"][0]) == d["ts_k"] and int(g["ts_R"][0]) == d["ts_R"]
    assert int(g["timeType"][0]) == d["time_type"] and int(g["MNtype"][0]) == d["MNtype"]
    sc = [float.fromhex(t) for t in g["scalars"]]
    want = [d["alpha"], d["tol"], d["kappa"], d["alphaD"], d["beta"], d["lD"], d["dD"], d["bD"], d["rDw"],
            d["rDwobs"], d["Lc"], d["Tc"], d["Hc"], d["l"], d["d"]]
    assert sc == [float(x) for x in want]
    for k, key in (("t", "t"), ("r", "r"), ("z", "z"), ("tD", "tD"), ("rD", "rD"), ("zD", "zD"), ("j0z", "j0z"),
                   ("timePar", "time_par"), ("MoenchGamma", "moench_gamma")):
        assert np.array_equal(hexv(g[k]), np.asarray(d[key], float)), k
    assert [int(x) for x in g["sv"][1:]] == [int(x) for x in d["sv"]]
    assert [int(x) for x in g["zLay"][1:]] == [int(x) for x in d["zLay"]]
    assert g["outfile"][0] == d["outfile"]


def test_fortran_edit_descriptors(cli):
    out = run(cli, "--format", "0.1", "1.9374742486", "-2.5e-7", "nan", "inf", "-inf", "0", "1e100", "123456.789").split("\n")
    assert out[0] == "[ 1.0000000E-01][ 1.000000000000000E-0001]"          # ES14.07E2 / ES24.15E4
    assert out[1] == "[ 1.9374742E+00][ 1.937474248600000E+0000]"          # SURVEY appendix B example
    assert out[2] == "[-2.5000000E-07][-2.500000000000000E-0007]"
    assert out[3] == "[           NaN][                     NaN]"
    assert out[4] == "[      Infinity][                Infinity]"
    assert out[5] == "[     -Infinity][               -Infinity]"
    assert out[6] == "[ 0.0000000E+00][ 0.000000000000000E+0000]"
    assert out[7] == "[**************][ 1.000000000000000E+0100]"          # 3-digit exponent overflows E2
    assert out[8] == "[ 1.2345679E+05][ 1.234567890000000E+0005]"


def test_headers_have_the_reference_layout(cli):
    h = run(cli, "hantush-input.dat", "--header-only").split("\n")
    assert len(h) - 1 == 20                         # plot-hantush-check.py:45 skips 20 rows
    assert h[0] == "# -*-auto-revert-*-" and h[1] == "# model, EP precision :: 1 Hantush, 8"
    assert h[2] == "# dimensionless?, timeseries?, piezometer? :: T T T "
    assert h[11] == "# deHoog M, alpha, tol :: 10 1.0000000E-08  1.0000000E-09 "
    assert h[-2] == "#" + "-" * 63 and h[-3].startswith("#     t_D              Hantush")
    c = run(cli, "theis-input.dat", "--header-only").split("\n")
    assert c[1] == "# model, EP :: 0 Theis, 8" and c[2] == "# dimensionless?, timeseries? :: T F "
    assert c[14].startswith("# num r locations, rlocs :: 30  7.5000000E-01 ")
    assert c[-2] == "#" + "-" * 76
    m = run(cli, "cape-cod-moench.in", "--header-only").split("\n")
    assert any(line.startswith("# characteristic head ::") for line in m)          # dimensional output
    assert m[-2].startswith("# Moench Delayed Yield decay coefficients (alpha):: 3 ")   # stdout side channel
    mn = run(cli, "mishra-neuman-malama.in", "--header-only")
    assert "# Mishra/Neuman ac,ak,psia,psik,b1 ::" in mn and "assumes ac=ak" in mn


def test_stale_format_deck_is_rejected_like_the_reference(cli, tmp_path):
    """A deck whose line 10 lacks MoenchM makes the list-directed read hit the '::' comment
    (SURVEY 3.3): the reference aborts, so does the reader."""
    lines = open(os.path.join(CFG, "hantush-input.dat")).read().split("\n")
    lines[9] = "0.0D0         :: beta only (stale format)"
    p = tmp_path / "stale.dat"
    p.write_text("\n".join(lines))
    r = subprocess.run([cli, str(p), "--dump"], capture_output=True, text=True, cwd=CFG)
    assert r.returncode != 0 and "bad integer" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hantush-input.dat", "cape-cod-moench.in", "theis-input.dat"])
def test_cli_end_to_end_writes_reference_layout(cli, name, tmp_path):
    d = deck.read_deck(os.path.join(CFG, name))
    pd = deck.params_dict(d)
    text = run(cli, name, "--stdout")
    rows = [l for l in text.split("\n") if l and not l.startswith("#")]
    stale = d["j0z"][d["sv"][0] - 1] / d["rD"][0]
    s, ds = ub.eval_grid(ub.Params(pd), d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"], ts_scale=stale)
    if d["timeseries"]:
        assert len(rows) == len(d["tD"])
        zo = d["zOrd"]
        for i, row in enumerate(rows):
            f, g = s[i, 0], ds[i, 0]
            if not d["piezometer"] and zo > 1:
                obs = (f[0] + 2.0 * f[1:zo].sum() + f[zo - 1]) / (2 * zo)
                der = (g[0] + 2.0 * g[1:zo].sum() + g[zo - 1]) / (2 * zo)
            else:
                obs, der = f[0], g[0]
            sc = 1.0 if d["dimless"] else d["Hc"]
            tcol = d["tD"][i] if d["dimless"] else d["t"][i]
            want = "%14.7E %s %s " % (tcol, fmt24(obs * sc), fmt24(der * sc))
            assert row == want, (i, row, want)
    else:
        assert len(rows) == len(d["rD"]) * len(d["zD"])
        k = 0
        for ir in range(len(d["rD"])):
            for iz in range(len(d["zD"])):
                want = "%14.7E %14.7E %s %s " % (d["zD"][iz], d["rD"][ir], fmt24(s[0, ir, iz]), fmt24(ds[0, ir, iz]))
                assert rows[k] == want, (k, rows[k], want)
                k += 1


def fmt24(x):
    if np.isnan(x):
        return "NaN".rjust(24)
    m, e = ("%.15E" % x).split("E")
    return (m + "E%s%04d" % (e[0], int(e[1:]))).rjust(24)
