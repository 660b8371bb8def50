"""dd_halo_push (csrc/dd_halo.cu): the direct-store halo exchange of the multi-process slab driver, exercised in ONE
process: three "ranks" are three streams of the same GPU with their own field and flag blocks (plain device
pointers stand in for the CUDA IPC mappings), so the kernels really run concurrently and handshake through the
flag words.  The multi-process path itself (IPC export / import over NVLink) is what `bench.py --gpus N` runs; its
`check.slabs_equal_whole_mesh` compares the result with an undecomposed run."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200"))

pytestmark = pytest.mark.gpu


def test_three_ranks_exchange_halos_by_peer_stores():
    import torch
    from _ddlib import Context
    world, rows, G, ld = 3, 40, 6, 1001  # odd pitch: the blocks of rows are not all 16-byte aligned
    streams = [torch.cuda.Stream() for _ in range(world)]
    ctxs = [Context(0, s.cuda_stream) for s in streams]
    lib = ctxs[0].lib
    flags = []
    for c in ctxs:
        f = C.c_void_p()
        c.check(lib.dd_halo_flags_create(c.handle, C.byref(f)), "flags")
        flags.append(f)
    # local layout of every rank: [G halo | rows owned | G halo] (edge ranks simply do not use one halo)
    own0, own1 = G, G + rows
    fields = [torch.zeros((rows + 2 * G, ld), dtype=torch.float64, device="cuda") for _ in range(world)]
    rng = np.random.default_rng(5)
    for seq in range(1, 6):
        truth = []
        for r in range(world):
            a = rng.normal(size=(rows, ld))
            truth.append(a)
            with torch.cuda.stream(streams[r]):
                fields[r][own0:own1].copy_(torch.from_numpy(a), non_blocking=False)
        torch.cuda.synchronize()
        for r in range(world):
            base = fields[r].data_ptr()
            up, down = r - 1, r + 1
            src_top = dst_up = src_bot = dst_down = upf = dnf = None
            if up >= 0:
                src_top = base + own0 * ld * 8
                dst_up = fields[up].data_ptr() + own1 * ld * 8
                upf = flags[up]
            if down < world:
                src_bot = base + (own1 - G) * ld * 8
                dst_down = fields[down].data_ptr() + (own0 - G) * ld * 8
                dnf = flags[down]
            ctxs[r].check(lib.dd_halo_push(ctxs[r].handle, src_top, dst_up, src_bot, dst_down, G * ld, flags[r], upf, dnf,
                                           seq), "push")
        torch.cuda.synchronize()
        for r in range(world):
            got = fields[r].cpu().numpy()
            assert np.array_equal(got[own0:own1], truth[r])
            if r > 0:
                assert np.array_equal(got[own0 - G:own0], truth[r - 1][rows - G:]), (seq, r)
            if r < world - 1:
                assert np.array_equal(got[own1:own1 + G], truth[r + 1][:G]), (seq, r)
            st = C.c_int(-1)
            ctxs[r].check(lib.dd_halo_status(ctxs[r].handle, flags[r], C.byref(st)), "status")
            assert st.value == 0
    for c, f in zip(ctxs, flags):
        c.check(lib.dd_halo_flags_destroy(c.handle, f), "destroy")
