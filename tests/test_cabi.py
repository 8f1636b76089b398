"""CPU: the C-ABI library builds/loads and exports every symbol include/barvae.h declares (no compute calls)."""
import os
import re

from gpu_util import ROOT, pkg


def test_library_exports_every_declared_symbol():
    lib = pkg("_lib")
    l = lib.load()
    hdr = open(os.path.join(ROOT, "include", "barvae.h")).read()
    declared = set(re.findall(r"\b(bvae_[a-z0-9_]+)\s*\(", hdr))
    bound = {name for name, _, _ in lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(l, name), name
    assert l.bvae_version() >= 100


def test_struct_layouts_match_header(tmp_path):
    """compile include/barvae.h with gcc and compare sizeof/offsetof of every descriptor field with the ctypes mirror"""
    import ctypes as C
    import subprocess
    lib = pkg("_lib")
    structs = {"bvae_conv_desc": lib.ConvDesc, "bvae_wgrad_desc": lib.WgradDesc, "bvae_nb_desc": lib.NbDesc,
               "bvae_pack_job": lib.PackJob, "bvae_bn_desc": lib.BnDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "barvae.h"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f[0], cname, f[0]))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f in cls._fields_:
            assert int(got["%s.%s" % (cname, f[0])]) == getattr(cls, f[0]).offset, (cname, f[0])


def test_no_cpu_fallback():
    """product modules refuse CPU tensors instead of silently computing elsewhere"""
    import pytest
    import torch
    Model = pkg("graph.model").Model
    enc = Model().encoder
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 1, 96, 60))
