"""CPU checks of the C-ABI boundary: the library loads and exports every symbol include/pyxu_b200.h
declares; the ctypes prototypes cover the same set; struct layouts agree with the C compiler."""
import ctypes as C
import os
import re
import subprocess
import tempfile

from conftest import ROOT
from pyxu_b200 import _build, _cabi as K

HEADER = os.path.join(ROOT, "include", "pyxu_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pxb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    _build.build_cuda()
    h = C.CDLL(K.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/pyxu_b200.h but not exported"
    assert set(names) == set(K.PROTOTYPES), set(names) ^ set(K.PROTOTYPES)
    assert K.lib().pxb_abi_version() == K.ABI_VERSION


def test_struct_layouts_match_the_c_compiler():
    prog = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "pyxu_b200.h"
    int main(void){
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(pxb_slab), sizeof(pxb_stencil_desc), sizeof(pxb_grad_desc),
             sizeof(pxb_prox_spec), sizeof(pxb_fterm), sizeof(pxb_pds_params), sizeof(pxb_stencil2d), sizeof(pxb_fista_step), sizeof(pxb_pad2d_desc),
             sizeof(pxb_stop_rule), sizeof(pxb_iter_ctl));
      printf("%zu %zu ", sizeof(pxb_peer), offsetof(pxb_peer, epoch));
      printf("%zu %zu ", offsetof(pxb_stop_rule, table), offsetof(pxb_iter_ctl, ticket));
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(pxb_stencil_desc, coef), offsetof(pxb_grad_desc, coef),
             offsetof(pxb_grad_desc, slab), offsetof(pxb_pds_params, lam), offsetof(pxb_stencil2d, coef), offsetof(pxb_stencil2d, add_period),
             offsetof(pxb_fista_step, norms), offsetof(pxb_stencil2d, origin), offsetof(pxb_pad2d_desc, mode));
      printf("%zu %zu %zu %zu %zu\n", sizeof(pxb_stencil3d), sizeof(pxb_stencil3d_dense), offsetof(pxb_stencil3d_dense, coef),
             offsetof(pxb_stencil3d_dense, slab), offsetof(pxb_stencil3d, slab));
      return 0; }
    """
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    extra = [C.sizeof(K.Peer), K.Peer.epoch.offset]
    sizes = [C.sizeof(s) for s in (K.Slab, K.StencilDesc, K.GradDesc, K.ProxSpec, K.FTerm, K.PdsParams, K.Stencil2D, K.FistaStep, K.Pad2D, K.StopRule, K.IterCtl)]
    offs = [K.StopRule.table.offset, K.IterCtl.ticket.offset, K.StencilDesc.coef.offset, K.GradDesc.coef.offset, K.GradDesc.slab.offset, K.PdsParams.lam.offset, K.Stencil2D.coef.offset,
            K.Stencil2D.add_period.offset, K.FistaStep.norms.offset, K.Stencil2D.origin.offset, K.Pad2D.mode.offset]
    st3 = [C.sizeof(K.Stencil3D), C.sizeof(K.Stencil3DDense), K.Stencil3DDense.coef.offset, K.Stencil3DDense.slab.offset, K.Stencil3D.slab.offset]
    assert [int(v) for v in out] == sizes + extra + offs + st3


def test_argument_errors_are_reported_without_a_gpu():
    lib = K.lib()
    d = K.StencilDesc()
    rc = lib.pxb_stencil_apply(C.byref(d), None, None, None)
    assert rc == -1 and b"null" in lib.pxb_last_error()
    s2 = K.Stencil2D()
    assert lib.pxb_stencil2d_apply(C.byref(s2), None, None, None) == -1
    assert lib.pxb_stencil2d_fista(C.byref(s2), None, 0, None, None) == -1
    one = (C.c_int64 * 3)(4, 4, 4)
    assert lib.pxb_stencil_axis0_apply(0, 1, one, None, 3, 5, (C.c_double * 3)(1, 2, 1), C.c_void_p(16), C.c_void_p(32), None) == -1  # center outside the kernel
    assert lib.pxb_pds_iter(0, None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.pxb_pds_iter_n(0, None, None, None, None, None, None, None, None, 4, None, None, None) == -1
    assert lib.pxb_pds_iter_p2p(0, None, None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.pxb_set_iter_path(7) == -1 and lib.pxb_set_iter_path(0) == 0
    assert lib.pxb_set_iter_modes(2) == -1
    g = K.GradDesc()
    g.ndir = 7
    rc = lib.pxb_gradient_apply(C.byref(g), C.c_void_p(8), C.c_void_p(16), None)
    assert rc in (-1, -3)
