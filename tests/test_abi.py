"""CPU tests: the C-ABI library loads and exports every symbol include/lmono.h declares (no
compute calls without a GPU); the oracle exports the lmono_cpu_* checker symbols."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lmono_[a-z0-9_]+)\s*\(", txt)))


def _exported(so):
    out = subprocess.run(["nm", "-D", "--defined-only", so], check=True, capture_output=True, text=True).stdout
    return {l.split()[-1] for l in out.splitlines() if l.strip()}


def test_cuda_library_builds_and_exports_the_abi():
    from lmono_b200 import api
    so = api.build()
    L = api.lib()
    names = _declared(os.path.join(ROOT, "include", "lmono.h"))
    assert len(names) >= 20
    exp = _exported(so)
    missing = [n for n in names if n not in exp]
    assert not missing, missing
    assert L.lmono_strerror(0) == b"ok"
    p = api.Params()
    L.lmono_default_params(__import__("ctypes").byref(p))
    assert p.scan_line == 64 and abs(p.mapping_plane_resolution - 0.8) < 1e-7


def test_cuda_library_has_sm100a_code_only():
    from lmono_b200 import api
    so = api.build()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_oracle_exports_checker_symbols(oracle):
    so = os.path.join(ROOT, "oracle", "liblmono_oracle.so")
    exp = _exported(so)
    names = _declared(os.path.join(ROOT, "oracle", "lmono_oracle.h"))
    missing = [n for n in names if n not in exp]
    assert not missing, missing


def test_product_code_never_touches_the_oracle():
    """The product path must not import / link / call anything under oracle/."""
    bad = []
    pkg = os.path.join(ROOT, "lmono_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle_lib|liblmono_oracle|lmono_cpu_|oracle/", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
    out = subprocess.run(["ldd", os.path.join(pkg, "csrc", "liblmono_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_every_pdl_launched_kernel_waits_for_its_predecessor():
    """Kernels launched with programmatic dependent launch (LM_LAUNCH_PDL, common.cuh) may start before the previous
    kernel of the stream has finished: each of them must execute lm_pdl_enter() (griddepcontrol.wait) as its first
    statement, or the transitive ordering of the step chain is lost."""
    csrc = os.path.join(ROOT, "lmono_b200", "csrc")
    src = {f: open(os.path.join(csrc, f)).read() for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))}
    launched = set()
    for txt in src.values():
        launched |= set(re.findall(r"LM_LAUNCH_PDL\(\s*(k_\w+)", txt))
    assert len(launched) >= 30, launched
    alltxt = "\n".join(src.values())
    for k in sorted(launched):
        defs = [m for m in re.finditer(r"__global__[^;{]*?\b%s\s*\(" % k, alltxt)]
        assert defs, k
        ok = False
        for m in defs:
            i = m.end(); d = 1
            while d > 0:
                d += {"(": 1, ")": -1}.get(alltxt[i], 0); i += 1
            body = alltxt[i:i + 80].lstrip()
            if body.startswith("{") and body[1:].lstrip().startswith("lm_pdl_enter();"):
                ok = True
        assert ok, f"{k} is launched with LM_LAUNCH_PDL but does not start with lm_pdl_enter()"
