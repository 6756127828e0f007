"""Where the time of ONE CTA of the dense kernels goes: per-phase clock64() stamps from a -DCWN_PHASE_TIMING build.

    python -c "from cwn_b200 import build; build.build_variant('phase', ['CWN_PHASE_TIMING'])"
    CWN_B200_LIB=$PWD/cwn_b200/csrc/libcwn_b200_phase.so python tools/phase_timing.py > gpurun_out/phase_timing.txt

For every grouped dense launch of one training step (warm), prints the CTA count, the kernel's span on the device
(%globaltimer of the first CTA start to the last stamp) and the median / max cycles between consecutive stamps.
"""
import collections
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cwn_b200 import _lib, fused  # noqa: E402
from cwn_b200.dist import FlatGradBucket  # noqa: E402
from cwn_b200.mp.molec_models import EmbedSparseCIN  # noqa: E402
from cwn_b200.optim import FlatAdam  # noqa: E402

MAX_CTAS = 4096


def main():
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    setter = lib.cwn_debug_set_phase_buffer
    setter.restype = ctypes.c_int
    setter.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(MAX_CTAS * 16, dtype=torch.int64, device=dev)
    torch.manual_seed(0)
    model = EmbedSparseCIN(**bench.MODEL_CFG).to(dev).train()
    bucket = FlatGradBucket(model)
    opt = FlatAdam(model, bucket, lr=1e-3)
    batch = bench.make_batches(1, 128, 1000)[0].to(dev)

    inputs = [batch.cochains[d].x for d in range(3)]

    def step():
        for d, x in enumerate(inputs):  # the embedding layers replace the integer features in place
            batch.cochains[d]._x = x
        loss = bench.l1(model(batch), batch.y)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    records = []
    orig = fused._launch

    def traced(fn_name, desc_type, descs):
        torch.cuda.synchronize()
        buf.zero_()
        torch.cuda.synchronize()
        _lib.check(setter(buf.data_ptr()))
        orig(fn_name, desc_type, descs)
        torch.cuda.synchronize()
        _lib.check(setter(None))
        records.append((fn_name, len(descs), buf.view(MAX_CTAS, 16).cpu().clone()))

    fused._launch = traced
    step()
    fused._launch = orig
    seen = collections.Counter()
    for name, n_prob, b in records:
        live = b[:, 0] != 0
        n = int(live.sum())
        if n == 0:
            continue
        key = (name, n_prob, n)
        seen[key] += 1
        if seen[key] > 1:
            continue
        b = b[live]
        t_start = b[:, 15]
        span_ns = int(t_start.max() - t_start.min())
        print(f'{name}  problems={n_prob}  CTAs={n}  CTA start spread {span_ns} ns')
        last = b[:, 0]
        for i in range(1, 15):
            have = b[:, i] != 0
            if not bool(have.any()):
                continue
            prev = torch.where(have, last, b[:, i])
            dt = (b[:, i] - prev)[have].float()
            tot = (b[:, i] - b[:, 0])[have].float()
            print(f'   phase {i - 1}->{i}: n={int(have.sum()):4d}  median {dt.median():8.0f}  max {dt.max():8.0f} cycles'
                  f'   (since CTA start: median {tot.median():8.0f}  max {tot.max():8.0f})')
            last = torch.where(have, b[:, i], last)


if __name__ == '__main__':
    main()
