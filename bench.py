"""bench.py — cells/sec of one training step of the ZINC-shaped ring-lifted CWN (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A "step" = one pass of the hot path over one batch of 128 synthetic ZINC-shaped complexes (V=23, E=25, rings
{6,6,5} -> 6 528 cells): EmbedSparseCIN (4 layers, hidden 64, edge embeddings, coboundary messages, BatchNorm,
sum readout) forward -> L1 loss -> backward -> gradient all-reduce (N>1) -> Adam step. The optimizer step is inside
the timed region although the metric is named fwd+bwd, so nothing is skipped.

  value : cells/s with the batch already resident in HBM (CSR plans are rebuilt every step: each step is a new
          batch), max over ranks, L2 flushed between steps
  e2e   : the same through the public API from HOST memory: ComplexBatch.to(device) (one pinned staging buffer
          per dtype) -> plans -> step -> loss.item()
  roofline : the dominant cwn kernel of the step, algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : the oracle (torch-only port of the reference path; the reference itself cannot be imported
          here: torch_scatter / torch_geometric are absent) on the host cores, bounded sample
  ragged : the same model on RAGGED batches (every step a different set of molecules, sizes varying around the ZINC
          means) padded to one fixed-capacity layout and replayed through one CUDA graph (cwn_b200.bucketed), next to
          the eager path on the same batches

    python bench.py --config ogb ...          # BASELINE.json configs[3]: OGBEmbedSparseCIN, molhiv-shaped, 2 layers,
                                              # mean readout, dropout 0.5 (reference exp/scripts/cwn-molhiv.sh)
    python bench.py --sweep-full [--gpus N]   # BASELINE.json configs[4]: per-adjacency sweep, independent shards
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# NCCL_DEBUG=VERSION makes NCCL print its banner on STDOUT at communicator creation; stdout of rank 0 must carry exactly
# one JSON line (the driver parses it), so the banner level is mapped to WARN (INFO and above are left to the user)
if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
    os.environ['NCCL_DEBUG'] = 'WARN'

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRELOAD_STEPS = 320  # untimed extra warm-up steps (~0.3 s) during which nvidia-smi samples the clocks under load
MODEL_CFG = dict(atom_types=28, bond_types=4, out_size=1, num_layers=4, hidden=64, dropout_rate=0.0, max_dim=2,
                 embed_edge=True, use_coboundaries=True, graph_norm='bn', readout='sum')
WORKLOAD = 'ZINC ring-lift (max_ring=6) EmbedSparseCIN 4-layer hidden=64 batch=128 (BASELINE.json configs[1])'
# BASELINE.json configs[3] (reference exp/scripts/cwn-molhiv.sh:3-33, mp/molec_models.py:281-350)
OGB_CFG = dict(out_size=1, num_layers=2, hidden=64, dropout_rate=0.5, indropout_rate=0.0, max_dim=2, jump_mode=None,
               nonlinearity='relu', readout='mean', final_readout='sum', apply_dropout_before='lin2',
               use_coboundaries=True, embed_edge=True, graph_norm='bn')
OGB_WORKLOAD = 'ogbg-molhiv-shaped ring-lift (max_ring=6) OGBEmbedSparseCIN 2-layer hidden=64 batch=128 dropout 0.5 ' \
               '(BASELINE.json configs[3])'
# ragged batches: ring count 1..4 (sizes 5..6), 4..14 chain atoms => V ~ 23, E ~ 25 on average (the ZINC means)
RAGGED_GEN = dict(ragged=True, ring_count_range=(1, 5), pendant_range=(4, 15))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='cwn', choices=['cwn', 'reference'])
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--pool', type=int, default=8, help='distinct pre-collated batches cycled through')
    ap.add_argument('--mode', default='auto', choices=['auto', 'eager', 'graph'])
    ap.add_argument('--no-sweep', action='store_true')
    ap.add_argument('--sweep-only', action='store_true', help='only the per-adjacency kernel sweep (for ncu)')
    ap.add_argument('--sweep-full', action='store_true', help='BASELINE config 5 grid, one JSON line per point')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-ragged', action='store_true', help='skip the ragged-batch leg (padded layout, one CUDA graph)')
    ap.add_argument('--config', default='zinc', choices=['zinc', 'ogb'], help="'ogb' = BASELINE.json configs[3]")
    return ap.parse_args()


def workload(args):
    """(model class name, constructor kwargs, workload string, generator kwargs, loss) of --config."""
    if args.config == 'ogb':
        return 'OGBEmbedSparseCIN', OGB_CFG, OGB_WORKLOAD, dict(ogb_features=True, num_pendant=8), bce
    return 'EmbedSparseCIN', MODEL_CFG, WORKLOAD, {}, l1


def bce(out, y):
    return torch.nn.functional.binary_cross_entropy_with_logits(out, (y.view(-1, 1) > 0).float())  # exp/train_utils.py:11


def l1(out, y):
    return torch.nn.functional.l1_loss(out, y.view(-1, 1))  # reference exp/train_utils.py:25-26


def make_batches(n_batches, batch_size, seed0, **gen):
    from cwn_b200.data import synthetic
    from cwn_b200.data.complex import ComplexBatch
    return [ComplexBatch.from_complex_list(synthetic.zinc_like_complexes(batch_size, seed=seed0 + i, **gen))
            for i in range(n_batches)]


def cells_of(batch):
    return sum(batch.cochains[d].num_cells for d in range(batch.dimension + 1))


# ------------------------------------------------------------------------------------------------ CPU arm
def host_threads():
    """Every host thread this process may use (torchrun pins OMP_NUM_THREADS=1: the CPU arm undoes that explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_steps(steps, warmup, batch_size, budget_s=None, args=None):
    """The oracle's forward + loss + backward + Adam on the host cores. Returns (cells/s, ms/step, steps done)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import cwn_oracle as O
    from cwn_b200.mp import molec_models
    name, cfg, _, gen, loss_fn = workload(args) if args is not None else ('EmbedSparseCIN', MODEL_CFG, None, {}, l1)
    oracle_fn = O.ogb_embed_sparse_cin if name == 'OGBEmbedSparseCIN' else O.embed_sparse_cin
    torch.set_num_threads(host_threads())
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in getattr(molec_models, name)(**cfg).state_dict().items()}
    leaves = []
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k and not k.endswith(('eps1', 'eps2')):
            v.requires_grad_(True)
            leaves.append(v)
    leaves = list({id(v): v for v in leaves}.values())
    opt = torch.optim.Adam(leaves, lr=1e-3)
    batches = make_batches(4, batch_size, seed0=1000, **gen)
    cells = cells_of(batches[0])
    done, t_total = 0, 0.0
    for i in range(warmup + steps):
        snap = O.Snapshot(batches[i % len(batches)])
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = oracle_fn(sd, cfg, snap, training=True)
        loss = loss_fn(out, snap.y)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            done += 1
            t_total += dt
            if budget_s is not None and t_total > budget_s:
                break
    ms = 1e3 * t_total / max(done, 1)
    return cells / (ms / 1e3), ms, done, cells


def run_reference(args, rank):
    """The reference arm: rank 0 alone, on ALL the host threads of the box (the other ranks exit without work). At
    N > 1 its line still describes the whole host: N GPUs are compared with the box's CPUs, not with one thread
    (torchrun's OMP_NUM_THREADS=1 is overridden in cpu_steps)."""
    if rank != 0:
        return
    value, ms, done, cells = cpu_steps(args.steps, args.warmup, args.batch, args=args)
    cores = torch.get_num_threads()
    wl = workload(args)[2]
    line = {
        'impl': 'reference', 'metric': 'cells/sec fwd+bwd ZINC ring-lifted CWN', 'value': value, 'unit': 'cells/s',
        'n_gpus': args.gpus, 'steps': done, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl, 'cells_per_step': cells, 'step': 'fwd+loss+bwd+adam',
                   'host': f'rank 0 only, {cores} host threads (the whole box), whatever --gpus says'},
        'cpu_baseline': {'value': value, 'unit': 'cells/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{done} full steps of the same workload (batch {args.batch}) through the '
                                   f'torch-only oracle; the reference cannot be imported (torch_scatter, '
                                   f'torch_geometric, torch_sparse absent)'},
        'e2e': {'value': value, 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """SM clock and throttle reasons of one GPU while the bench runs. Two sources feed the same sample list:
    * NVML polled in-process every ~4 ms (nvidia-ml-py) — fine enough to land samples INSIDE the ~20 ms timed region;
    * `nvidia-smi -lms 20` as a subprocess — the recipe's own clocks line, and the fallback if NVML is unavailable.
    Every failure of a source is swallowed: the sampler must never take the bench down."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index, nvml=None):
        self.samples, self.proc, self.gpu = [], None, gpu_index  # samples: (row, host arrival time, source)
        self._nvml, self._stop = nvml, False

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-i', str(self.gpu), '-lms', os.environ.get('CWN_BENCH_SMI_MS', '20')],
                                         stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        try:
            if self._nvml is None:
                import pynvml
                self._nvml = pynvml
            self._nvml.nvmlInit()
            handle = self._nvml.nvmlDeviceGetHandleByIndex(int(self.gpu))
            threading.Thread(target=self._poll_nvml, args=(handle,), daemon=True).start()
        except Exception:  # no NVML here (or no permission): nvidia-smi alone
            self._nvml = None

    def _pump(self):
        try:
            for line in self.proc.stdout:
                self.samples.append(([c.strip() for c in line.split(',')], time.perf_counter(), 'nvidia-smi'))
        except Exception:
            pass

    def _poll_nvml(self, handle):
        n = self._nvml
        try:
            sm_max = n.nvmlDeviceGetMaxClockInfo(handle, n.NVML_CLOCK_SM)
            get_reasons = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                n.nvmlDeviceGetCurrentClocksThrottleReasons
            flag = lambda mask, bit: 'Active' if (mask & bit) else 'Not Active'  # noqa: E731
            while not self._stop:
                sm = n.nvmlDeviceGetClockInfo(handle, n.NVML_CLOCK_SM)
                mask = int(get_reasons(handle))
                row = [str(self.gpu), str(int(sm)), str(int(sm_max)), '', hex(mask),
                       flag(mask, n.nvmlClocksThrottleReasonHwSlowdown),
                       flag(mask, n.nvmlClocksThrottleReasonHwThermalSlowdown),
                       flag(mask, n.nvmlClocksThrottleReasonSwThermalSlowdown),
                       flag(mask, n.nvmlClocksThrottleReasonSwPowerCap)]
                self.samples.append((row, time.perf_counter(), 'nvml'))
                time.sleep(0.004)
        except Exception:
            pass

    def wait_first(self, timeout_s=3.0):
        """nvidia-smi needs ~0.1-0.5 s before its first line: without this a 20 ms timed region ends unsampled."""
        t0 = time.perf_counter()
        while (self.proc is not None or self._nvml is not None) and not self.samples \
                and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def stop(self, t_from=None, t_to=None, t_timed=None):
        """Statistics over the samples that arrived in [t_from, t_to] (the period the GPU was under this bench's load);
        `t_timed` = (start, end) of the timed region, to report how many samples fell inside it."""
        self._stop = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        if self.proc is None and self._nvml is None and not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi and NVML unavailable'], 'samples': 0}
        keep = [(r, t, src) for r, t, src in list(self.samples)
                if (t_from is None or t >= t_from) and (t_to is None or t <= t_to)]
        in_timed = sum(1 for _, t, _ in keep if t_timed is not None and t_timed[0] <= t <= t_timed[1])
        rows = [r for r, _, _ in keep]
        sm = sorted(int(r[1]) for r in rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'samples_in_timed_region': in_timed,
                'sources': sorted({src for _, _, src in keep}),
                'window': 'NVML every ~4 ms + nvidia-smi every 20 ms, from the pre-load steps through the timed region '
                          'to the end of the e2e loop (all of it this workload under load; the timed region itself '
                          'lasts ~20 ms)'}


# ------------------------------------------------------------------------------------------------ GPU arm
def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return json.load(open(path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md: 6.65 TB/s)'


# Every timed launch of the sweeps sits behind a ~60 us spin kernel on its stream and the cost of an empty event pair is
# subtracted (ops.KernelProfile): a 10k-cell pass runs for 3-5 us, less than the host needs to issue it, so without the pad
# the event pair measures the launch path (8-25 us read as "0.03 of the peak"), not the kernel.
SWEEP_PAD = 120_000


class CleanL2Flush(object):
    """Evict the previous iteration from L2 and leave the cache CLEAN: a 256 MB write (larger than the 126 MB L2) followed
    by a 256 MB read of another buffer. After a write-only flush the L2 is full of dirty lines whose write-back (up to
    126 MB = ~20 us of DRAM time) is charged to whatever kernel runs next — a quarter of a 100 us pass that is being
    compared with a copy bandwidth measured over gigabytes."""

    def __init__(self, dev):
        self.w = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
        self.r = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def zero_(self):
        self.w.zero_()
        if not os.environ.get('CWN_BENCH_DIRTY_FLUSH'):  # A/B: what the write-only flush costs the next kernel
            self.r.sum()


def kernel_sweep(dev):
    """BASELINE config 5, one point: block-diagonal edge-upper adjacency at 1M edges/dim, F=64 — the regime where
    the HBM roofline is the bound. Algorithmic bytes per SURVEY 8(d)."""
    from cwn_b200 import ops
    from cwn_b200.data import synthetic
    peak, _ = peaks()
    out = []
    flush = CleanL2Flush(dev)
    for kind, F in [('edge_up', 64), ('edge_boundary', 64), ('edge_up', 256), ('edge_up', 16)]:
        index, cob, n_src, n_dst, n_cob = synthetic.tiled_adjacency(kind, 40_000)
        index = index.to(dev)
        x = torch.randn(n_src, F, device=dev)
        ops.gather_scatter(x, index, n_dst)  # builds the plan
        ms = []
        for _ in range(5):
            flush.zero_()
            with ops.KernelProfile(pad_cycles=SWEEP_PAD) as prof:
                ops.gather_scatter(x, index, n_dst)
            rec = prof.summary()['csr_gather_reduce']
            ms.append(rec['ms'])
        t = sorted(ms)[len(ms) // 2]
        gbs = rec['bytes'] / (t * 1e-3) / 1e9
        out.append({'kernel': 'csr_gather_reduce', 'adjacency': kind, 'F': F, 'cells': n_dst, 'messages': index.size(1),
                    'ms': t, 'algorithmic_GBps': gbs, 'frac_of_peak': gbs / peak})
        if cob is not None and F in (16, 64):
            cob = cob.to(dev)
            P, Q = torch.randn(n_src, F, device=dev, requires_grad=True), torch.randn(n_cob, F, device=dev, requires_grad=True)
            o = ops.cob_pass(P, Q, index, cob, n_dst)
            o.sum().backward()
            g = torch.randn_like(o)
            for name in ('csr_cob_fwd', 'csr_cob_bwd'):
                ms = []
                for _ in range(5):
                    P.grad = Q.grad = None
                    flush.zero_()
                    with ops.KernelProfile(pad_cycles=SWEEP_PAD) as prof:
                        o = ops.cob_pass(P, Q, index, cob, n_dst)
                        o.backward(g)
                    rec = prof.summary()[name]
                    ms.append(rec['ms'] / rec['launches'])
                t = sorted(ms)[len(ms) // 2]
                gbs = rec['bytes'] / rec['launches'] / (t * 1e-3) / 1e9
                out.append({'kernel': name, 'adjacency': kind, 'F': F, 'cells': n_dst, 'messages': index.size(1),
                            'ms': t, 'algorithmic_GBps': gbs, 'frac_of_peak': gbs / peak})
    return out


def kernel_sweep_full(dev, rank=0, world=1):
    """BASELINE config 5: N in {1e4,1e5,1e6} cells/dim x F in {16,64,256} x the four adjacency types, block-diagonal
    ZINC-like layout (+ one uniform-random stress point per F); forward, transposed (backward) and coboundary passes.
    One JSON object per point on stdout."""
    from cwn_b200 import ops
    from cwn_b200.data import synthetic
    peak, _ = peaks()
    flush = CleanL2Flush(dev)

    def timed(fn, name, reps=5):
        ms = []
        for _ in range(reps):
            flush.zero_()
            with ops.KernelProfile(pad_cycles=SWEEP_PAD) as prof:
                fn()
            rec = prof.summary()[name]
            ms.append(rec['ms'] / rec['launches'])
        t = sorted(ms)[len(ms) // 2]
        return t, rec['bytes'] / rec['launches']

    units_per = {'edge_boundary': 25, 'ring_boundary': 3, 'vertex_up': 23, 'edge_up': 25}
    points = [(k, n, 'block-diagonal') for k in units_per for n in (10_000, 100_000, 1_000_000)] + \
             [('random', 1_000_000, 'uniform-random')]
    for kind, n_cells, layout in points:
        if kind == 'random':
            index, cob, n_src, n_dst, n_cob = synthetic.random_adjacency(n_cells, n_cells, 3_200_000, n_cells // 8)
        else:
            index, cob, n_src, n_dst, n_cob = synthetic.tiled_adjacency(kind, max(1, n_cells // units_per[kind]))
        index = index.to(dev)
        cob = cob.to(dev) if cob is not None else None
        for F in (16, 64, 256):
            x = torch.randn(n_src, F, device=dev, requires_grad=True)
            out = ops.gather_scatter(x, index, n_dst)
            g = torch.randn_like(out)
            out.backward(g)  # builds the transposed plan
            rows = []
            t, b = timed(lambda: ops.gather_scatter(x.detach(), index, n_dst), 'csr_gather_reduce')
            rows.append(('identity_fwd', t, b))

            def bwd():
                x.grad = None
                ops.gather_scatter(x, index, n_dst).backward(g)
            tb = []
            for _ in range(3):
                flush.zero_()
                with ops.KernelProfile(pad_cycles=SWEEP_PAD) as prof:
                    bwd()
                prof.summary()  # (synchronises; measures the empty-pair overhead)
                recs = [r for r in prof.records if r[0] == 'csr_gather_reduce']
                tb.append(max(recs[1][2].elapsed_time(recs[1][3]) - prof.overhead_ms, 0.0))
            rows.append(('identity_bwd', sorted(tb)[1], recs[1][1]))
            if cob is not None:
                P = torch.randn(n_src, F, device=dev, requires_grad=True)
                Q = torch.randn(n_cob, F, device=dev, requires_grad=True)
                o = ops.cob_pass(P, Q, index, cob, n_dst)
                o.backward(g)

                def both():
                    P.grad = Q.grad = None
                    ops.cob_pass(P, Q, index, cob, n_dst).backward(g)
                t, b = timed(both, 'csr_cob_fwd')
                rows.append(('cob_fwd', t, b))
                t, b = timed(both, 'csr_cob_bwd')
                rows.append(('cob_bwd (avg of dP, dQ)', t, b))
            for name, t, b in rows:
                gbs = b / (t * 1e-3) / 1e9
                print(json.dumps({'rank': rank, 'n_gpus': world, 'pass': name, 'adjacency': kind, 'layout': layout, 'cells': n_dst,
                                  'messages': index.size(1), 'F': F, 'us': round(1e3 * t, 2),
                                  'algorithmic_MB': round(b / 1e6, 3), 'GBps': round(gbs, 1),
                                  'frac_of_measured_peak': round(gbs / peak, 4)}), flush=True)
            del x, out, g


def run_cwn(args, rank, world, local_rank):
    import torch.distributed as dist
    from cwn_b200 import _lib, ops
    from cwn_b200.dist import FlatGradBucket, broadcast_parameters
    from cwn_b200.optim import FlatAdam
    from cwn_b200.mp import molec_models
    model_name, model_cfg, workload_name, gen, loss_fn = workload(args)

    if not torch.cuda.is_available():
        raise RuntimeError('bench.py: no CUDA device; the cwn_b200 path is CUDA-only (use --impl reference for the CPU arm)')
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    _lib.load()
    if args.sweep_only:
        print(json.dumps({'kernel_sweep': kernel_sweep(dev)}), flush=True)
        return
    if args.sweep_full:  # independent shards: every rank sweeps its own GPU, lines carry the rank (weak scaling)
        kernel_sweep_full(dev, rank, world)
        return
    torch.manual_seed(0)
    model = getattr(molec_models, model_name)(**model_cfg).to(dev).train()
    broadcast_parameters(model)
    bucket = None
    if world > 1 and os.environ.get('CWN_BENCH_DP_SYMM', '1') != '0':
        # gradients in symmetric memory: the all-reduce is fused into the Adam kernel over NVLink peer memory (no NCCL call)
        try:
            from cwn_b200.dist import SymmetricGradBucket
            bucket = SymmetricGradBucket(model)
            bucket.self_test()  # one fused launch on known data at THIS world size, before anything depends on it
        except Exception as exc:  # noqa: BLE001 — reported, never silent: config.allreduce says which path ran
            print(f'bench.py: symmetric memory unavailable ({type(exc).__name__}: {exc}); NCCL all-reduce',
                  file=sys.stderr, flush=True)
    if bucket is None:
        bucket = FlatGradBucket(model)
    opt = FlatAdam(model, bucket, lr=1e-3)  # one launch; also clears the gradient bucket for the next step
    fused_dp = getattr(opt, 'fuses_allreduce', False)

    host_batches = [b.pack_(pin_memory=True) for b in make_batches(args.pool, args.batch, seed0=1000 + 100 * rank, **gen)]
    cells = cells_of(host_batches[0])
    dev_batches = [b.to(dev) for b in make_batches(args.pool, args.batch, seed0=1000 + 100 * rank, **gen)]
    index_tensors = [[t for d in range(3) for t in (b.cochains[d].upper_index, b.cochains[d].boundary_index,
                                                    b.cochains[d].batch)] for b in dev_batches]
    inputs = [[b.cochains[d].x for d in range(3)] for b in dev_batches]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2

    def step(batch):
        out = model(batch)
        loss = loss_fn(out, batch.y)
        loss.backward()
        if not fused_dp:
            bucket.all_reduce()
        opt.step()
        return loss

    def eager_resident_step(i):
        b = dev_batches[i % args.pool]
        ops.clear_plan_cache(*index_tensors[i % args.pool])  # every step is a new batch: plans are rebuilt
        for d, x in enumerate(inputs[i % args.pool]):         # the forward overwrites cochain.x (set_xs): restore
            b.cochains[d]._x = x
        return step(b)

    # ---- whole-step CUDA graph (plans + fwd + loss + bwd [+ adam]) over static packed buffers
    mode, captured, launches_per_step = 'eager', None, None
    if args.mode in ('auto', 'graph'):
        from cwn_b200.graph import CapturedStep
        try:
            static = make_batches(1, args.batch, seed0=999, **gen)[0].to(dev)
            dp_graph = world > 1 and os.environ.get('CWN_BENCH_DP_GRAPH', '1') != '0'
            captured = None
            if dp_graph:  # the all-reduce captured into the step graph (one launch per step)
                try:
                    captured = CapturedStep(model, loss_fn, bucket, opt, optimizer_in_graph=True, allreduce_in_graph=True)
                    captured.capture(static)
                except Exception as exc:  # noqa: BLE001
                    print(f'bench.py: capturing the NCCL all-reduce failed ({type(exc).__name__}: {exc}); three-piece step',
                          file=sys.stderr, flush=True)
                    captured, dp_graph = None, False
                    torch.cuda.synchronize()
                    static = make_batches(1, args.batch, seed0=999, **gen)[0].to(dev)
            if captured is None:
                captured = CapturedStep(model, loss_fn, bucket, opt, optimizer_in_graph=(world == 1))
                captured.capture(static)
            l0 = _lib.launch_count()
            captured._body()  # one eager pass of the captured body: counts the cwn kernels one replay launches
            launches_per_step = _lib.launch_count() - l0
            mode = 'cuda-graph'
        except Exception as exc:  # noqa: BLE001 — reported in the JSON line, never silent
            if args.mode == 'graph':
                raise
            print(f'bench.py: CUDA-graph capture failed ({type(exc).__name__}: {exc}); running eagerly',
                  file=sys.stderr, flush=True)
            captured, mode = None, f'eager (graph capture failed: {type(exc).__name__})'
            torch.cuda.synchronize()

    def resident_step(i):
        if captured is not None:
            return captured.run(dev_batches[i % args.pool])
        return eager_resident_step(i)

    def e2e_step(i):
        if captured is not None:
            captured.run(host_batches[i % args.pool])  # pinned host -> static device buffers, one copy per dtype
            return float(captured.loss.item()), host_batches[i % args.pool].packed_nbytes
        db = make_host_copy(host_batches[i % args.pool]).to(dev)
        return float(step(db).item()), db._h2d_bytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (the clock sampler runs from here to the end of the timed region: the region itself lasts tens of ms)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        clocks.wait_first()
    t_load0 = time.perf_counter()
    for i in range(max(args.warmup, 3)):
        resident_step(i)
    # pre-load: keep the GPU busy with the same step for ~0.3 s so that the clocks are ramped and sampled under load
    # before the (tens of ms long) timed region starts
    # (a FIXED number of steps: every rank must issue the same count, each step holds an all-reduce)
    for i in range(PRELOAD_STEPS):
        resident_step(i)
        if i % 32 == 31:
            torch.cuda.synchronize()
    barrier()

    # ---- timed region: device-resident inputs
    l0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_timed0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        resident_step(i)
        ev[i][1].record()
    barrier()
    t_timed1 = time.perf_counter()
    launches = _lib.launch_count() - l0 if captured is None else launches_per_step * args.steps
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = cells * world / (ms_step / 1e3)

    # ---- e2e: host batch (pinned) -> device -> step -> loss back on the host, every step
    for i in range(3):
        e2e_step(i)
    barrier()
    e2e_ms, h2d = 0.0, 0
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss_value, h2d = e2e_step(i)
        e2e_ms += 1e3 * (time.perf_counter() - t0)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = cells * world / (float(t.item()) / args.steps / 1e3)
    clock_info = clocks.stop(t_load0, time.perf_counter(), (t_timed0, t_timed1)) if rank == 0 else None

    # ---- e2e with collation inside the timed region: the dataset lives in HBM (PackedComplexDataset), a step is
    #      ids -> GPU collation straight into the graph's static buffers -> step -> loss.item()
    collated = None
    if captured is not None:
        from cwn_b200.data import synthetic
        from cwn_b200.data.packed import PackedComplexDataset
        n_ds = 8 * args.batch
        ds = PackedComplexDataset(synthetic.zinc_like_complexes(n_ds, seed=5000 + rank, **gen), max_dim=2, device=dev)
        perm = torch.randperm(n_ds, generator=torch.Generator().manual_seed(rank)).tolist()
        pick = lambda i: perm[(i * args.batch) % n_ds:(i * args.batch) % n_ds + args.batch]  # noqa: E731
        for i in range(3):
            ds.collate(pick(i), out=captured.static)
            captured.run()
        barrier()
        # Every step: table H2D + collate kernel + graph replay + that step's loss read back. The host half of the NEXT
        # collation (numpy: segment sizes, prefix sums) runs while the GPU executes the current step — after the step
        # has been launched, before its loss is read.
        pinned = torch.empty(1, dtype=torch.float32).pin_memory()
        done = torch.cuda.Event()
        c_ms, prep = 0.0, ds.prepare(pick(0))
        for i in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ds.launch(prep, out=captured.static)
            captured.run()
            pinned.copy_(captured.loss.detach().reshape(1), non_blocking=True)
            done.record()
            prep = ds.prepare(pick(i + 1))
            done.synchronize()
            float(pinned[0])
            c_ms += 1e3 * (time.perf_counter() - t0)
        t = torch.tensor([c_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        collated = {'value': cells * world / (float(t.item()) / args.steps / 1e3), 'unit': 'cells/s',
                    'what': 'dataset resident in HBM; per step: 128 ids -> GPU collation (one kernel, written into the '
                            'graph\'s static buffers; the segment table of the NEXT batch is prepared by the host while the GPU runs the step) -> '
                            'step -> loss.item(); the only host->device traffic is the segment table', 'h2d_bytes_per_step': int(ds._table_bytes),
                    'd2h_bytes_per_step': 4}

    # ---- ragged batches (every step a different set of molecules): padded to ONE layout, replayed through one graph
    ragged = None
    if not args.no_ragged and world == 1 and args.config == 'zinc':
        ragged = ragged_leg(args, dev, model, bucket, opt, flush)

    # ---- per-kernel roofline of the step (instrumented eager re-run of the same steps; every rank takes part because
    #      the step contains the gradient all-reduce)
    with ops.KernelProfile(pad_cycles=120_000) as prof:  # ~60 us of spin before each launch: see KernelProfile
        for i in range(min(args.steps, 5)):
            flush.zero_()
            eager_resident_step(i)
    n_prof = min(args.steps, 5)
    summary = prof.summary()
    barrier()
    if rank != 0:
        return
    peak, peak_src = peaks()
    # (the fused all-reduce + Adam kernel waits for its peers INSIDE the launch: in this eager, padded re-run the ranks
    # drift apart by milliseconds, so its event time is peer wait, not kernel time — reported under a name that says so
    # and never taken for the dominant kernel)
    if 'allreduce_adam_step' in summary:
        summary['allreduce_adam_step (incl. waiting for the slowest rank of the eager re-run)'] = summary.pop('allreduce_adam_step')
    dom = max((k for k in summary if k != 'csr_plan_build' and not k.startswith('allreduce_adam_step')),
              key=lambda k: summary[k]['ms'])
    rec = summary[dom]
    achieved = rec['bytes'] / (rec['ms'] * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), if any
    traffic, tpath = None, os.path.join(ROOT, 'profiles', 'r2_ncu_dram_traffic.json')
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, 'profiles', 'r1_ncu_dram_traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom, {}).get('dram_bytes_per_launch')
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'as_contraction': tensor_roofline(dom, rec),
                'launches_per_step': rec['launches'] / n_prof,
                'avg_launch_us': 1e3 * rec['ms'] / rec['launches'],
                'algorithmic_bytes_per_launch': rec['bytes'] / rec['launches'],
                'kernel_ms_per_step': {k: v['ms'] / n_prof for k, v in summary.items()},
                'timed_in': 'instrumented eager re-run of the same steps: CUDA events on the launching stream around '
                            'every C-ABI launch, each preceded by a ~60 us spin kernel so that the event pair is queued '
                            'before the GPU reaches it and brackets device time only (an eager step is host-bound); the median '
                            'cost of an empty event pair is subtracted',
                'event_pair_overhead_us': 1e3 * prof.overhead_ms,
                'note': 'at batch 128 every adjacency pass moves ~1-2 MB (L2-resident): the step is launch/latency '
                        'bound, see kernel_sweep for the HBM-bound regime'}
    line = {
        'metric': 'cells/sec fwd+bwd ZINC ring-lifted CWN', 'value': value, 'unit': 'cells/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name, 'cells_per_step_per_gpu': cells, 'global_batch': args.batch * world,
                   'parallelism': f'dp{world}', 'step': 'plans+fwd+loss+bwd+allreduce+adam', 'mode': mode,
                   'allreduce': ('fused with Adam in one kernel over NVLink peer memory (symmetric memory, two-shot; no NCCL '
                                 'call on the step)' if fused_dp else
                                 'captured in the step graph' if (captured is not None and captured.allreduce_in_graph) else
                                 'between two graphs' if world > 1 else 'none (1 GPU)'), 'clock_preload_steps': PRELOAD_STEPS,
                   'l2': 'flushed (256 MB write) between timed steps', 'last_loss': loss_value},
        'clocks': clock_info, 'gpu_launches': int(launches),
        'e2e': {'value': e2e_value, 'unit': 'cells/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 4},
        'roofline': roofline,
    }
    if collated is not None:
        line['e2e_gpu_collation'] = collated
    if ragged is not None:
        line['ragged'] = ragged
    if not args.no_sweep:
        line['kernel_sweep'] = kernel_sweep(dev)
    if not args.no_cpu_baseline:
        v, ms, done, _ = cpu_steps(10, 2, args.batch, budget_s=20.0, args=args)
        line['cpu_baseline'] = {'value': v, 'unit': 'cells/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                'ms_per_step': ms,
                                'sample': f'{done} full steps of the same workload through the torch-only oracle'}
    print(json.dumps(line), flush=True)


def tensor_roofline(name, rec):
    """The dense kernels are contractions, not streams: their algorithmic flops (2 n K h per product, fp32-equivalent;
    the 3xTF32 split issues four TF32 products for each) against the TF32 tensor peak = half the MEASURED bf16 peak."""
    flops = rec.get('flops', 0)
    if not flops:
        return None
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    bf16 = json.load(open(path))['bf16_tflops'] if os.path.exists(path) else 1590.0
    achieved = flops / (rec['ms'] * 1e-3) / 1e12
    return {'bound': 'tensor', 'kernel': name, 'achieved': achieved, 'unit': 'TFLOP/s (fp32-equivalent)', 'peak': bf16 / 2,
            'peak_source': 'TF32 dense = MEASURED_PEAKS.json bf16_tflops / 2', 'frac': achieved / (bf16 / 2),
            'tf32_products_per_fp32_product': 4,
            'note': 'latency-bound at this size: ~100-200 CTAs of 64 x 64 x 64..128 tiles, one wave'}


def ragged_leg(args, dev, model, bucket, opt, flush):
    """Ragged batches through cwn_b200.bucketed.BucketedStep (one graph, padded layout) and, for comparison, through the
    eager path on their own layouts. Runs after the uniform legs on the same model (weights keep training)."""
    from cwn_b200.bucketed import BucketedStep, Capacity, masked_l1
    from cwn_b200.data import synthetic
    from cwn_b200.data.complex import ComplexBatch
    n_pool = max(args.pool, 8)
    pool = synthetic.zinc_like_complexes(n_pool * args.batch, seed=7000, **RAGGED_GEN)
    lists = [pool[i * args.batch:(i + 1) * args.batch] for i in range(n_pool)]
    cap = Capacity.from_dataset(pool, args.batch)
    step = BucketedStep(model, masked_l1, bucket, opt, capacity=cap, optimizer_in_graph=True)
    step.capture(lists[0])
    host = [step.pad(lst) for lst in lists]                      # padded + packed + pinned
    resident = [step.pad(lst, pin_memory=False).to(dev) for lst in lists]
    real_cells = [sum(int(b.cochains[d].live_rows) for d in range(3)) for b in host]
    layouts = len({tuple(int(b.cochains[d].live_rows) for d in range(3)) for b in host})
    for i in range(5):
        step.run(resident[i % n_pool])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        step.run(resident[i % n_pool])
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    cells = sum(real_cells[i % n_pool] for i in range(args.steps)) / args.steps
    # end to end: padded pinned host batch -> static buffers -> replay -> loss.item()
    for i in range(3):
        step.run(host[i % n_pool])
    torch.cuda.synchronize()
    e2e = 0.0
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step.run(host[i % n_pool])
        float(step.loss.item())
        e2e += 1e3 * (time.perf_counter() - t0)
    e2e /= args.steps
    # the eager path on the same (unpadded) batches: what a ragged batch cost before
    eager_batches = [ComplexBatch.from_complex_list(lst).to(dev) for lst in lists]
    eager_inputs = [[b.cochains[d].x for d in range(3)] for b in eager_batches]

    def eager(i):
        b = eager_batches[i % n_pool]
        for d, x in enumerate(eager_inputs[i % n_pool]):
            b.cochains[d]._x = x
        loss = l1(model(b), b.y)
        loss.backward()
        opt.step()
    n_eager = min(args.steps, 10)
    for i in range(n_pool):  # every layout once: the caching allocator must have seen each batch's sizes before timing
        eager(i)             # (two warm-up steps left cudaMallocs inside the timed loop: 69 ms instead of 12 ms)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_eager):
        eager(i)
    torch.cuda.synchronize()
    eager_ms = 1e3 * (time.perf_counter() - t0) / n_eager
    return {'value': cells / (ms / 1e3), 'unit': 'cells/s', 'ms_per_step': ms, 'real_cells_per_step': cells,
            'e2e': {'value': cells / (e2e / 1e3), 'unit': 'cells/s', 'h2d_bytes_per_step': int(host[0].packed_nbytes),
                    'd2h_bytes_per_step': 4},
            'distinct_layouts_in_pool': layouts, 'capacity': repr(cap), 'padded_cells_per_step': sum(cap.cells),
            'eager_ms_per_step': eager_ms, 'eager_value': cells / (eager_ms / 1e3),
            'what': f'{n_pool} ragged batches of {args.batch} molecules (ring count 1..4, 4..14 chain atoms: ZINC-like '
                    'means) padded to one fixed-capacity layout and replayed through ONE CUDA graph; BatchNorm and '
                    'the loss see the real cells only (cwn_b200/bucketed.py); cells/s counts real cells; the eager '
                    'number is the same model on the same batches without padding or graph (host wall clock)'}


def make_host_copy(batch):
    """Fresh host-side ComplexBatch sharing the (CPU) tensors of `batch` (`.to()` rebinds attributes in place)."""
    import copy
    new = copy.copy(batch)
    new.cochains = {d: copy.copy(c) for d, c in batch.cochains.items()}
    return new


def main():
    args = parse()
    if os.environ.get('CWN_BENCH_WATCHDOG'):  # dump every thread's stack if the run is still alive after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['CWN_BENCH_WATCHDOG']), exit=True)
    from cwn_b200.dist import init_from_env
    if args.impl == 'reference':
        rank = int(os.environ.get('RANK', '0'))
        run_reference(args, rank)
        return
    rank, world, local_rank = init_from_env()
    run_cwn(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
