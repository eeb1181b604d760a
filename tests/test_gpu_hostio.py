"""-m gpu: host<->device staging at the boundary (successor of the reference's data_conversion.py) and the
host-buffer form of the transport stencil.  Behavioural spec: the reference's bridge test writes 11 into a
2x2 array_float from Python and Fortran checks it (test/py_ftn_interface/data/runtime_fortran.f90:32-43,
fortran_program.f90:27-33) -- i.e. pointer round trip and Fortran (column-major) layout."""
import cffi
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fortran_pointer_round_trip_and_layout():
    from b200stencil.hostio import HostDeviceConversion

    ffi = cffi.FFI()
    conv = HostDeviceConversion()
    ni, nj, nk = 4, 3, 2
    # Fortran array A(ni,nj,nk), column-major: element (i,j,k) at i + ni*(j + nj*k)
    buf = ffi.new("double[]", ni * nj * nk)
    for k in range(nk):
        for j in range(nj):
            for i in range(ni):
                buf[i + ni * (j + nj * k)] = 100 * i + 10 * j + k
    dev = conv.fortran_to_device(buf, [ni, nj, nk])
    conv.sync()
    assert dev.stride(0) == 1 and tuple(dev.shape) == (ni, nj, nk)
    assert float(dev[3, 2, 1]) == 321.0 and float(dev[1, 0, 1]) == 101.0
    # KAT of the reference bridge test: Python writes 11 everywhere, the caller's memory sees it
    out = ffi.new("float[]", 2 * 2)
    dev_out = conv.fortran_to_device(out, [2, 2])
    conv.sync()
    dev_out.fill_(11.0)
    conv.device_to_fortran(dev_out, out)
    assert [out[i] for i in range(4)] == [11.0] * 4
    # fp64 arrays come back whole (the reference copied 4*size bytes for every dtype, data_conversion.py:95)
    dev.mul_(2.0)
    conv.device_to_fortran(dev, buf)
    assert buf[ni * nj * nk - 1] == 2 * (100 * (ni - 1) + 10 * (nj - 1) + (nk - 1))
    with pytest.raises(TypeError):
        conv.device_to_fortran(dev.float(), buf)  # no casting at the boundary


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_fv_tp2d_host_pipeline_matches_resident(dtype):
    from b200stencil import fields, stencils
    from b200stencil.hostio import FvTp2dHost

    ni, nj, nk, nb = 40, 24, 3, 5
    pipe = FvTp2dHost(ni, nj, nk, dtype)
    host = pipe.host_fields(nb)
    g = torch.Generator().manual_seed(11)
    for name in pipe.NAMES:
        lo, hi = (0.5, 1.5) if name in ("q", "rarea") else (-0.9, 0.9)
        host[name].uniform_(lo, hi, generator=g)
    assert host["q"].is_pinned()
    pipe(host)
    assert pipe.h2d_bytes > 0 and pipe.d2h_bytes > 0
    for b in range(nb):
        d = {n: fields.from_numpy(host[n][b].numpy()) for n in pipe.NAMES}
        out = fields.zeros((ni, nj, nk), dtype)
        stencils.fv_tp2d(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
        assert torch.equal(out.cpu(), host["q_out"][b])
