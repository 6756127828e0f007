"""Fork/join execution of independent branches on concurrent CUDA streams.

Within a layer the cochain dimensions are independent of each other, and inside a cochain the upper-adjacency
branch and the boundary branch are independent until `combine_nn` (reference `mp/layers.py:184-199, 333-342` runs
them one after the other). Each of their kernels covers only a few thousand rows — a fraction of the 148 SMs — so
the branches are issued on side streams that fork from and re-join the current stream. Under CUDA-graph capture
(`cwn_b200/graph.py`) the fork/join becomes graph edges: the replayed DAG runs the branches concurrently with no
host involvement. Autograd replays each backward node on the stream of its forward op, so the backward pass gets
the same concurrency.

Memory safety relies on the fork/join discipline: every side stream waits for its parent before starting and the
parent waits for all side streams before continuing, so a block freed after the join cannot be handed to a kernel
that runs before its last reader.

Set CWN_B200_STREAMS=0 to run everything on the current stream.
"""
import os

import torch

ENABLED = os.environ.get('CWN_B200_STREAMS', '1') != '0'
_pools = {}


def _side_streams(parent: torch.cuda.Stream, n: int):
    key = (parent.device, parent.cuda_stream)
    pool = _pools.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=parent.device))
    return pool[:n]


def run_concurrently(thunks, device=None):
    """Run the zero-argument callables `thunks` as concurrent branches; returns their results in order."""
    thunks = list(thunks)
    if (not ENABLED or len(thunks) < 2 or device is None or torch.device(device).type != 'cuda'):
        return [t() for t in thunks]
    parent = torch.cuda.current_stream(device)
    streams = _side_streams(parent, len(thunks))
    outs = []
    for s, t in zip(streams, thunks):
        s.wait_stream(parent)
        with torch.cuda.stream(s):
            outs.append(t())
    for s in streams:
        parent.wait_stream(s)
    return outs


_aux = {}


def aux_stream(parent: torch.cuda.Stream) -> torch.cuda.Stream:
    """A dedicated side stream per parent for work whose JOIN is deferred (forked in forward, awaited by an event in
    backward): kept apart from the fork/join pool so that an unrelated join does not wait for it."""
    key = (parent.device, parent.cuda_stream)
    s = _aux.get(key)
    if s is None:
        s = _aux[key] = torch.cuda.Stream(device=parent.device)
    return s


# ---------------------------------------------------------------------------------------------- deferred tail work
_tails, _pending = {}, {}


def run_deferred(fn, keepalive, device):
    """Run `fn()` (kernel launches) on a dedicated tail stream forked from the current one, and join it only when the
    running BACKWARD pass ends (autograd's end-of-pass callback). For work nothing in the rest of the backward pass
    reads — the ordered sums of the weight-gradient partials into `.grad`: issued in line they sat on the critical
    path of every layer (two launches of 5-15 us each per layer), as a parallel branch they cost nothing.

    `keepalive`: every tensor the launches touch; the references are dropped after the join, so the caching allocator
    cannot hand their memory to a later kernel of the parent stream while the tail still reads it. Outside a backward
    pass (or with CWN_B200_STREAMS=0) `fn` simply runs in line."""
    if not ENABLED or torch.device(device).type != 'cuda':
        fn()
        return
    engine = torch.autograd.Variable._execution_engine
    parent = torch.cuda.current_stream(device)
    key = (parent.device, parent.cuda_stream)
    state = _pending.setdefault(key, {'keep': [], 'queued': False})
    if not state['queued']:
        def join(device=parent.device, key=key, state=state):
            # (autograd runs end-of-pass callbacks on the stream backward() was called from, after it has been
            # synchronised with the streams the pass used: that is the stream whoever reads `.grad` next will be on)
            torch.cuda.current_stream(device).wait_stream(_tails[key])
            state['keep'] = []
            state['queued'] = False
        try:
            engine.queue_callback(join)
        except RuntimeError:  # not inside a backward pass
            fn()
            return
        state['queued'] = True
    tail = _tails.get(key)
    if tail is None:
        tail = _tails[key] = torch.cuda.Stream(device=parent.device)
    tail.wait_stream(parent)
    with torch.cuda.stream(tail):
        fn()
    state['keep'].append(keepalive)
