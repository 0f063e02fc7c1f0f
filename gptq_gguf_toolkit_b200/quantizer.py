"""Block-sequential quantisation driver -- host-side mirror of the reference's
quant/gptq/src/quantizer.py::Quantizer (same constructor arguments, same quantize(quant_config) call, same
<save_dir>/<module name>/data.pth schema, quantizer.py:120-128,206-214,267-275).

What is different (B200-first, results unchanged):
  * the per-layer work runs in libgq (include/gq.h) instead of torch loops;
  * layers of a block that see the same input tensor (q/k/v, gate/up) share ONE Hessian and ONE U, and,
    when they also share the q_type, are quantised as one stacked matrix in one kernel launch (rows of a
    GPTQ problem are independent given U, so the outputs are bit-identical to separate calls);
  * calibration sequences of equal length are pushed through a block in batches (`calibration_batch_size`)
    instead of one at a time (GPTQ.update already weights a batch by its number of sequences, gptq.py:110-111);
  * multi-GPU (torch.distributed, NCCL): every rank holds its slice of the calibration sequences (as in the
    reference, quant.py:177-179), Hessians are all-reduced (gptq.py:131-132), then -- unlike the reference,
    where rank 0 quantises alone -- every rank quantises a ROW SLICE of each projection and the results are
    all-gathered; the data.pth files are written by the ranks in turn (`spread_emission`), or by rank 0 alone.
"""
from __future__ import annotations

import contextlib
import os
import threading
import queue
from typing import Any, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .gptq import GPTQ, HessianAccumulator
from .model_utils import ForwardInterrupt, InputCollector, LINEAR_LAYERS, maybe_first_element, select_layers, to
from .quant_utils import GGML_QUANT_SIZES, GGMLQuantizationType


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _rank() -> int:
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def _world() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


_BCAST_GROUPS: dict = {}      # (chain slot, world size) -> process group, see Quantizer._bcast_group


class PhaseTimer:
    """CUDA-event phase timing without synchronising inside the run; totals() syncs once at the end."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled and torch.cuda.is_available()
        self.spans: Dict[str, list] = {}

    def span(self, name: str):
        timer = self

        class _Span:
            def __enter__(self_inner):
                if timer.enabled:
                    self_inner.a = torch.cuda.Event(enable_timing=True)
                    self_inner.b = torch.cuda.Event(enable_timing=True)
                    self_inner.a.record()
                return self_inner

            def __exit__(self_inner, *exc):
                if timer.enabled:
                    self_inner.b.record()
                    timer.spans.setdefault(name, []).append((self_inner.a, self_inner.b))
                return False

        return _Span()

    def totals(self) -> Dict[str, float]:
        if not self.enabled:
            return {}
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) / 1e3 for k, v in self.spans.items()}


class _AsyncSaver:
    """torch.save on a worker thread so that file I/O overlaps GPU work (the reference saves inline)."""

    def __init__(self):
        self.q: "queue.Queue" = queue.Queue(maxsize=16)
        self.err: Optional[BaseException] = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            obj, path, ready = item
            try:
                if ready is not None:
                    ready.synchronize()     # the D2H copies were enqueued before this event
                os.makedirs(os.path.dirname(path), exist_ok=True)
                torch.save(obj, path)
            except BaseException as e:  # noqa: BLE001
                self.err = e

    def submit(self, obj, path, ready):
        self.q.put((obj, path, ready))

    def close(self):
        self.q.put(None)
        self.t.join()
        if self.err is not None:
            raise self.err


class Quantizer:
    def __init__(self, model: nn.Module, data_loader: Iterable, quantizable_modules: str,
                 quantizer_kwargs: Dict[str, Any], pre_block_modules: List[str], post_block_modules: List[str],
                 block_modules: str, save_dir: Optional[str], quant_non_block_modules: bool = False,
                 device: Optional[torch.device] = None, cpu_offload_modules: bool = False,
                 cpu_offload_activations: bool = False, verbose: bool = False,
                 # ---- additions over the reference ----
                 calibration_batch_size: int = 8, share_hessians: bool = True, keep_results: bool = False,
                 save_packed: bool = True, timer: Optional[PhaseTimer] = None, early_exit_pass1: bool = True,
                 overlap_prepare: bool = True, defer_last_layer: bool = True, fused_forward_ops: bool = True,
                 early_prepare: bool = True, rtn_native_arith: bool = True, shard_prepare: bool = True,
                 concurrent_groups: bool = True, spread_emission: bool = True) -> None:
        self.model = model
        self.data_loader = data_loader
        self.quantizable_modules = quantizable_modules
        self.quantizer_kwargs = dict(quantizer_kwargs)
        self.pre_block_modules = pre_block_modules
        self.post_block_modules = post_block_modules
        self.block_modules = block_modules
        self.device = device
        self.cpu_offload_modules = cpu_offload_modules
        self.cpu_offload_activations = cpu_offload_activations
        self.quant_non_block_modules = quant_non_block_modules
        self.verbose = verbose
        self.save_dir = save_dir
        self.calibration_batch_size = max(1, int(calibration_batch_size))
        self.share_hessians = share_hessians
        self.keep_results = keep_results
        self.save_packed = save_packed
        self.early_exit_pass1 = early_exit_pass1
        self.overlap_prepare = overlap_prepare
        self.defer_last_layer = defer_last_layer
        self.fused_forward_ops = fused_forward_ops
        self.early_prepare = early_prepare
        # embed_tokens / lm_head of a 16-bit model: scale search in the weight's own arithmetic like the reference
        # (quantizer.py:303-305 -> gq_rtn_quantize_native); False = weights widened to fp32
        self.rtn_native_arith = rtn_native_arith
        # several ranks: every Cholesky chain (gq_prepare) of a block runs on ONE owner rank and U is broadcast, instead of
        # every rank running all of them (the reference factors on every rank too, gptq.py:305-324 is not rank-guarded)
        self.shard_prepare = shard_prepare
        # the column loops of a block's groups (q/k/v, o, gate/up: different layers, nothing in common) run concurrently, each
        # on its group's side stream right behind its Cholesky chain, instead of one after the other on the main stream: the
        # serial panel -> update -> panel chains of the groups interleave and their CTA waves (192 + 128 + 896 CTAs of 32 rows on
        # 148 SMs for Llama-3-8B) fill up together
        self.concurrent_groups = concurrent_groups
        # several ranks: every rank holds every result after the all-gathers, so the device -> host copies and the data.pth
        # writes of the modules are dealt out round-robin over the ranks (N PCIe links and N writer threads instead of rank 0's);
        # the files are the same ones the reference's rank 0 writes (quantizer.py:120-128, 206-214, 267-275).  False: rank 0 only.
        self.spread_emission = spread_emission
        self._emit_counter = 0
        self.fused_installed: List[str] = []
        self._split_ok: Optional[bool] = None
        self._side_streams: list = []
        self._gather_stream = None
        self._gather_stream_used = False
        self._mask_flags: list = []
        self.timer = timer or PhaseTimer(False)
        self.results: Dict[str, Dict[str, Any]] = {}    # module name -> data.pth dict (+ "packed"), if keep_results
        self.non_invertible: List[str] = []
        self._saver: Optional[_AsyncSaver] = None

    # -------------------------------------------------------------------------------------------
    def _log(self, msg: str):
        if self.verbose and _rank() == 0:
            print(msg, flush=True)

    def _get_submodule(self, module_name: str):
        return self.model.get_submodule(module_name)

    def _create_handle(self, layer, hessian: Optional[HessianAccumulator] = None):
        return GPTQ(layer, hessian=hessian, **self.quantizer_kwargs)          # quantizer.py:239-240

    # -------------------------------------------------------------------------------------------
    def _emit(self, name: str, q_type, five, packed):
        """data.pth schema of the reference (quantizer.py:267-275) + optional 'packed' GGUF bytes; written by ONE rank: rank 0,
        or with `spread_emission` the module's turn in a round-robin over the ranks (every rank calls this for every module in
        the same order, so the counter agrees everywhere)."""
        owner = (self._emit_counter % _world()) if self.spread_emission else 0
        self._emit_counter += 1
        if _rank() != owner:
            return
        if self.save_dir is None and not self.keep_results:
            return
        qweight, d, sq, dmin, zq = five

        def host(t):   # pinned destination => the copy is truly asynchronous (torch caches pinned blocks)
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=t.is_cuda)
            buf.copy_(t, non_blocking=True)
            return buf

        with self.timer.span("save_d2h"):
            obj = {
                "q_type": int(q_type),
                "qweight": host(qweight),
                "super_group_scale": host(d),
                "super_group_zero": host(dmin),
                "group_scale_quant": host(sq),
                "group_zero_quant": host(zq),
            }
            if self.save_packed and packed is not None:
                obj["packed"] = host(packed)
            ready = None
            if qweight.is_cuda:
                ready = torch.cuda.Event()
                ready.record()
        if self.keep_results:
            self.results[name] = obj
        if self.save_dir is not None:
            self._saver.submit(obj, os.path.join(self.save_dir, name, "data.pth"), ready)

    # -------------------------------------------------------------------------------------------
    def _quant_non_block_module(self, name: str, module: nn.Module, quant_config):
        """quantizer.py:94-128 / 181-214 / 278-330: RTN K-quant of embed_tokens / lm_head (default Q6_K)."""
        q_type = quant_config.get(name.split(".")[-1], GGMLQuantizationType.Q6_K)
        kw = self.quantizer_kwargs
        w = module.weight.data
        rtn_kw = dict(packed=self.save_packed, wdeq_dtype=w.dtype, **({"native_arith": True} if self.rtn_native_arith else {}))
        world, rank = _world(), _rank()
        if world > 1:
            # rows of an RTN problem are independent: every rank quantises a 32-row-aligned slice (128256 x 4096 for Llama-3:
            # 41 ms on one GPU in the reference's bf16 arithmetic) and the slices are all-gathered, like the GPTQ layers
            total = w.shape[0]
            per = -(-(-(-total // world)) // 32) * 32
            lo, hi = min(rank * per, total), min((rank + 1) * per, total)
            wl = torch.zeros(per, w.shape[1], dtype=w.dtype, device=w.device)
            wl[: hi - lo] = w[lo:hi]
            outs = ops.rtn_quantize(wl, int(q_type), kw.get("rmin", -1.0), kw.get("rdelta", 0.1), kw.get("nstep", 20), **rtn_kw)
            full = [None if t is None else torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in outs]
            with self.timer.span("allgather"):
                for g, t in zip(full, outs):
                    if t is not None:
                        dist.all_gather_into_tensor(g, t.contiguous())
            qweight, d, sq, dmin, zq, packed, wdeq = (None if g is None else g[:total] for g in full)
        else:
            qweight, d, sq, dmin, zq, packed, wdeq = ops.rtn_quantize(
                w.contiguous(), int(q_type), kw.get("rmin", -1.0), kw.get("rdelta", 0.1), kw.get("nstep", 20), **rtn_kw)
        module.weight.data = wdeq
        self._emit(name, q_type, (qweight, d, sq, dmin, zq), packed)

    # -------------------------------------------------------------------------------------------
    def _prepare_hooks_and_handles(self, layers: Dict[str, nn.Module], quant_config=None, n_batches: int = 0):
        """quantizer.py:222-237, plus (a) Hessian sharing: layers that receive the very same input tensor object
        during a forward are attached to one accumulator, which is updated once per forward; (b) pass-1 early
        exit: the hooks are forward PRE-hooks (they see the same `inp[0]` as the reference's forward hooks), the
        first forward of a block records the order in which the hooked layers fire, and every later forward is
        interrupted (ForwardInterrupt) right after the last layer's Hessian update -- the reference discards the
        output of pass 1 (quantizer.py:150-151), so the last projection's GEMM and the block tail are dead work;
        (c) early prepare: a Hessian is complete as soon as its update of the LAST calibration batch has been enqueued,
        so the group's all-reduce, working copy and Cholesky chain (side stream) are started right there, inside the
        hook, and run underneath the rest of that forward and the other layers' Hessian updates instead of after pass 1
        (same kernels on the same inputs: bit-neutral)."""
        handles: Dict[str, GPTQ] = {}
        hooks = {}
        seen: list = []      # [(input tensor, accumulator)] of the forward in flight (refs keep addresses unique)
        state = {"grouped": False, "order": [], "last": None, "fired": 0, "batch": 0, "early": {}, "simple": False}

        def maybe_start_chain(h):
            # state["simple"]: every hooked layer fired exactly once in the recorded forward (one update per batch)
            if not (self.early_prepare and self.overlap_prepare and quant_config is not None and state["grouped"]
                    and state["simple"] and n_batches > 1 and state["batch"] == n_batches - 1 and h.layer.weight.is_cuda):
                return
            groups = self._group_names(handles)
            if len(groups) < 2 or id(h.hessian) in state["early"]:
                return
            for gi, names in enumerate(groups):
                if handles[names[0]].hessian is h.hessian:
                    state["early"][id(h.hessian)] = self._plan_group(gi, names, handles, quant_config, True)
                    return

        def make_hook(name):
            def _hook(_, inp):
                x = inp[0]
                h = handles[name]
                state["fired"] += 1
                if not state["grouped"]:
                    state["order"].append(name)
                if not self.share_hessians:
                    h.update(x)
                    maybe_start_chain(h)
                else:
                    for t, acc in seen:
                        if t is x:
                            if h.hessian is not acc:
                                if state["grouped"]:
                                    raise RuntimeError(f"inconsistent input sharing for {name}")
                                h.hessian.users -= 1
                                h.hessian = acc
                                acc.users += 1
                            break
                    else:
                        h.update(x)
                        seen.append((x, h.hessian))
                        maybe_start_chain(h)
                if state["last"] == name and state["fired"] == len(handles):
                    raise ForwardInterrupt
            return _hook

        for layer_name, layer in layers.items():
            handles[layer_name] = self._create_handle(layer)
            hooks[layer_name] = layer.register_forward_pre_hook(make_hook(layer_name))

        def end_of_forward():
            seen.clear()
            if not state["grouped"]:
                state["grouped"] = True
                order = state["order"]
                state["simple"] = bool(order) and len(order) == len(handles) == len(set(order))
                # early exit only when every hooked layer fired exactly once in the recorded forward
                if self.early_exit_pass1 and state["simple"]:
                    state["last"] = order[-1]
            state["fired"] = 0
            state["batch"] += 1

        return handles, hooks, end_of_forward, state

    @staticmethod
    def _group_names(handles: Dict[str, GPTQ]) -> List[List[str]]:
        """Layer names per shared HessianAccumulator, in module order."""
        groups: Dict[int, List[str]] = {}
        for name, h in handles.items():
            groups.setdefault(id(h.hessian), []).append(name)
        return list(groups.values())

    def _chain_owner(self, gi: int, handles: Dict[str, GPTQ]) -> int:
        """Owner rank of group gi's Cholesky chain: the chains of a block (cost ~ d_col^3) are dealt out greedily, most
        expensive first, to the least loaded rank -- counting ranks from the LAST one down, because rank 0 also moves the
        results to the host and writes the files.  Deterministic, so every rank computes the same assignment."""
        world = _world()
        groups = self._group_names(handles)
        cost = [float(handles[g[0]].d_col) ** 3 for g in groups]
        load = [0.0] * world
        owner = [0] * len(groups)
        for g in sorted(range(len(groups)), key=lambda i: (-cost[i], i)):
            r = min(range(world - 1, -1, -1), key=lambda k: (load[k], -k))
            owner[g] = r
            load[r] += cost[g]
        return owner[gi]

    def _bcast_group(self, gi: int):
        """One process group (NCCL communicator) per chain slot: a broadcast of U waits for its chain on the side stream,
        and collectives of one communicator execute in issue order -- on the default group the short chains' results and
        the all-gathers would queue behind the longest chain."""
        key = (gi, _world())      # process-wide: creating a communicator costs ~0.5 s, a Quantizer may be created per run
        if key not in _BCAST_GROUPS:
            _BCAST_GROUPS[key] = dist.new_group(ranks=list(range(_world())))
        return _BCAST_GROUPS[key]

    def _prepare_or_receive(self, gi: int, handles, H, W, rel_damp, side, slot):
        """gq_prepare on the chain's owner rank + broadcast of (U, not-PD flag) on the chain's stream; plain ops.prepare with
        one rank or shard_prepare off."""
        if not (_dist_on() and self.shard_prepare):
            return ops.prepare(H, W, rel_damp, stream=side, slot=slot)
        owner = self._chain_owner(gi, handles)
        grp = self._bcast_group(gi)
        if _rank() == owner:
            U, flag = ops.prepare(H, W, rel_damp, stream=side, slot=slot)
        else:
            U = torch.empty(H.shape[0], H.shape[0], dtype=torch.float32, device=H.device)
            flag = torch.zeros(1, dtype=torch.int32, device=H.device)
            if side is not None:
                side.wait_stream(torch.cuda.current_stream(H.device))
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            with self.timer.span("bcast_u"):
                dist.broadcast(U, src=owner, group=grp)
                dist.broadcast(flag, src=owner, group=grp)
        return U, flag

    def _plan_group(self, gi: int, names: List[str], handles: Dict[str, GPTQ], quant_config, overlap: bool):
        """Phase A of one group (quantizer.py:242-255 up to the factor): all-reduce of H, stacked fp32 working copy,
        dead-channel fix, Cholesky chain (on side stream `gi` when `overlap`).  Returns the plan tuple consumed by
        _launch_group / _finish_group."""
        kw = self.quantizer_kwargs
        hs = [handles[n] for n in names]
        acc = hs[0].hessian
        rows = [h.d_row for h in hs]
        side = self._side_stream(gi) if overlap else None
        # fp32 working copy of all members, stacked row-wise (gptq.py:138)
        W = torch.cat([h.layer.weight.data.float().flatten(1, -1) for h in hs], dim=0).contiguous()
        # With several ranks and a side stream the Hessian all-reduce and the dead-channel fix run on the side stream too
        # (their only consumer is the Cholesky chain behind them), so that the forward in flight on the main stream -- this
        # runs inside a hook of the last calibration batch -- is not held up by the collective.
        pre = side if (side is not None and _dist_on()) else None
        if pre is not None:
            pre.wait_stream(torch.cuda.current_stream(W.device))
        with (torch.cuda.stream(pre) if pre is not None else contextlib.nullcontext()):
            with self.timer.span("allreduce"):
                acc.all_reduce()                                               # gptq.py:131-132
            with self.timer.span("prepare_host"):
                ops.pre_step(acc.H, W)                                         # gptq.py:134-141
                if len(hs) > 1:
                    masks = torch.stack([(w == 0).all(dim=0) for w in W.split(rows, dim=0)])
                    self._mask_flags.append((names, (masks != masks[0:1]).any()))   # checked at the end (no host sync here)
        if pre is not None and kw.get("act_order", False):
            torch.cuda.current_stream(W.device).wait_stream(pre)      # act_order reads diag(H) on the main stream below
        with self.timer.span("prepare_host"):
            # act_order (gptq.py:209-216): the loop runs on W[:, perm] with the factor of H[perm][:, perm]; Q3_K
            # members ignore it (gptq.py:204-206) and need the plain factor
            q3 = [quant_config.get(n.split(".")[-1], GGMLQuantizationType.Q4_K) == GGMLQuantizationType.Q3_K for n in names]
            perm = U_perm = U = None
            not_pd = []          # device flags (one per factorisation), read only at the end
            if kw.get("act_order", False) and not all(q3):
                perm = torch.argsort(torch.diag(acc.H), descending=True)
                Hp = acc.H.index_select(0, perm).index_select(1, perm).contiguous()
                Wp = W.index_select(1, perm).contiguous()
                if side is not None:      # temporaries of this function that the side stream still has to read
                    Hp.record_stream(side)
                    Wp.record_stream(side)
                U_perm, flag = ops.prepare(Hp, Wp, kw.get("rel_damp", 1e-2), stream=side, slot=1 + gi if overlap else 0)
                not_pd.append(flag)
            if perm is None or any(q3):
                U, flag = self._prepare_or_receive(gi, handles, acc.H, W, kw.get("rel_damp", 1e-2), side,
                                                   1 + gi if overlap else 0)                               # gptq.py:305-324
                not_pd.append(flag)
            done = None
            if side is not None:
                done = torch.cuda.Event()
                done.record(side)
        return (names, hs, rows, W, U, not_pd, done, side, perm, U_perm)

    # -------------------------------------------------------------------------------------------
    def _quant_group(self, handles: Dict[str, GPTQ], quant_config, defer: Optional[str] = None, early: Optional[dict] = None):
        """quantizer.py:242-275.  Handles are processed per shared accumulator; same-q_type members of a group are
        stacked row-wise into one launch, and with several ranks each rank takes a row slice.
        defer: name of a layer whose (single-layer) group is NOT finished here: its Cholesky chain AND its column loop
        are enqueued on the group's side stream and a callable is returned that, when called later on the main
        stream, waits for them, swaps the layer's weight and emits the result (see _deferred_tail_plan)."""
        groups = self._group_names(handles)
        rank, world = _rank(), _world()
        on_gpu = next(iter(handles.values())).layer.weight.is_cuda
        overlap = bool(self.overlap_prepare) and on_gpu and len(groups) > 1
        staged = overlap and self.overlap_prepare == "staged"
        main = torch.cuda.current_stream() if on_gpu else None
        # ---- phase A: per distinct Hessian: all-reduce, fp32 working copy, dead-channel fix, U = chol(inv(H)).
        # With `overlap_prepare` every group's Cholesky chain runs on its own side stream (own workspace slot), so
        # the latency-bound chains of the 4 groups of a block run concurrently.  True / "eager": the main stream
        # waits only for the U it is about to consume, so later chains also overlap the column-loop kernels of the
        # groups before them; "staged": all chains first, then the column loops.  Groups whose chain was already started
        # from the forward hook of the last calibration batch (`early`) are taken as they are.
        early = early or {}
        plans = []
        for gi, names in enumerate(groups):
            plan = early.get(id(handles[names[0]].hessian))
            plans.append(plan if plan is not None else self._plan_group(gi, names, handles, quant_config, overlap))
        deferred = None
        if defer is not None:
            hit = [pl for pl in plans if pl[0] == [defer]]
            if overlap and len(hit) == 1 and hit[0][7] is not None:
                deferred = hit[0]
                plans = [pl for pl in plans if pl is not deferred]
        # ---- phase B: the column loops, in module order, on the main stream
        if staged:      # "staged": the column loops start only when every chain has finished (no SM contention)
            for plan in plans:
                if plan[6] is not None:
                    main.wait_event(plan[6])
        # With several ranks the all-gathers of a group's results run on a communication stream (_sharded_gptq), so the
        # next group's column loop does not wait for them: launch everything first, then wait once and finish in order.
        launched_all = []
        conc = bool(self.concurrent_groups) and overlap and not staged and len(plans) > 1 and all(pl[7] is not None for pl in plans)
        conc_events = []
        for plan in plans:
            if conc:       # column loop on the group's own stream, right behind its chain (bit-neutral: same kernels, same inputs)
                launched_all.append((plan, self._launch_group(plan, quant_config, rank, world, plan[7])))
                ev = torch.cuda.Event()
                ev.record(plan[7])
                conc_events.append(ev)
                continue
            if plan[6] is not None and not staged:
                main.wait_event(plan[6])
            launched_all.append((plan, self._launch_group(plan, quant_config, rank, world, None)))
        for ev in conc_events:
            main.wait_event(ev)
        launched = ready = None
        if deferred is not None:
            # the deferred group: column loop on ITS side stream, right behind its Cholesky chain
            launched = self._launch_group(deferred, quant_config, rank, world, deferred[7])
            ready = torch.cuda.Event()
            ready.record(deferred[7])
        if self._gather_stream_used:
            main.wait_stream(self._gather_stream)
            self._gather_stream_used = False
        for plan, l in launched_all:
            self._finish_group(plan, l)
        if deferred is None:
            return None

        def finish():
            main.wait_event(ready)
            self._finish_group(deferred, launched)
        return finish

    def _launch_group(self, plan, quant_config, rank, world, stream):
        """Column loops of one group (one launch per q_type present in it); results are not consumed yet."""
        names, hs, rows, W, U = plan[:5]
        q_types = [quant_config.get(n.split(".")[-1], GGMLQuantizationType.Q4_K) for n in names]  # quantizer.py:249
        dtype = hs[0].layer.weight.dtype
        by_type: Dict[int, List[int]] = {}          # sub-groups of equal q_type, keeping module order
        for i, qt in enumerate(q_types):
            by_type.setdefault(int(qt), []).append(i)
        offs = [0]
        for r in rows:
            offs.append(offs[-1] + r)
        launched = []
        perm, U_perm = plan[8], plan[9]
        static = bool(self.quantizer_kwargs.get("static_groups", False))
        for qt, idxs in by_type.items():
            self._log(f"Quantizing {[names[i] for i in idxs]} with {GGMLQuantizationType(qt).name}.")
            Wg = W if len(idxs) == len(hs) else torch.cat([W[offs[i]:offs[i + 1]] for i in idxs], 0).contiguous()
            if qt == int(GGMLQuantizationType.Q3_K) or perm is None:
                outs = self._sharded_gptq(Wg, U, qt, dtype, rank, world, stream, static and qt != int(GGMLQuantizationType.Q3_K), None)
            else:
                outs = self._sharded_gptq(Wg, U_perm, qt, dtype, rank, world, stream, True, perm)
            launched.append((qt, idxs, outs))
        return launched

    def _finish_group(self, plan, launched):
        """Write the dequantised weights back into the layers (quantizer.py:257-264), emit data.pth, reset handles."""
        names, hs, rows, W, U, not_pd = plan[:6]
        for qt, idxs, outs in launched:
            qweight, d, sq, dmin, zq, packed, wdeq = outs
            r0 = 0
            for i in idxs:
                r1 = r0 + rows[i]
                h = hs[i]
                wnew = wdeq[r0:r1].clone() if len(idxs) > 1 else wdeq[r0:r1]
                h.layer.weight.data = wnew.reshape(h.W_shape)      # convolutions keep their kernel shape
                self._emit(names[i], qt, (qweight[r0:r1], d[r0:r1], sq[r0:r1], dmin[r0:r1], zq[r0:r1]),
                           packed[r0:r1] if packed is not None else None)
                r0 = r1
        self._not_pd_flags.append((names, not_pd))
        for h in hs:
            h.reset()                                                          # quantizer.py:265

    def _side_stream(self, i: int):
        """Side streams for the Cholesky chains.  They run at HIGH priority: their kernels are small and latency-bound,
        the column-loop kernels on the main stream have hundreds of long-lived CTAs; with equal priorities a chain's
        next kernel queues behind all pending column-loop CTAs and the chain starves."""
        while len(self._side_streams) <= i:
            lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
            self._side_streams.append(torch.cuda.Stream(priority=hi))
        return self._side_streams[i]

    def _sharded_gptq(self, W, U, qt, dtype, rank, world, stream=None, static_groups=False, perm=None):
        kw = self.quantizer_kwargs
        slot = None
        if stream is not None and stream in self._side_streams:
            slot = 100 + self._side_streams.index(stream)      # one fast-mode scratch buffer per concurrent column loop
        args = dict(block_size=kw.get("block_size", 128) or W.shape[1], rmin=kw.get("rmin", -1.0),
                    rdelta=kw.get("rdelta", 0.1), nstep=kw.get("nstep", 20), packed=True, wdeq_dtype=dtype,
                    mode={"exact": 0, "fast": 1, "exact_left": 2, "exact_right": 3}[kw.get("mode", "exact")], static_groups=static_groups, perm=perm,
                    ws_slot=slot)
        if world == 1:
            return ops.gptq_quantize(W, U, qt, stream=stream, **args)[:7]
        # Row slice of this rank (in units of the kernel's 32-row CTA tile), then one all-gather per result tensor.
        # Every buffer is allocated on the CURRENT stream; with `stream` (the deferred group's side stream) the column
        # loop and the all-gathers are enqueued there -- torch's NCCL work waits for / is waited on by the stream that
        # is current at the call, and every rank issues the collectives in the same host order.
        total = W.shape[0]
        per = -(-total // world)
        per = -(-per // 32) * 32
        lo, hi = min(rank * per, total), min((rank + 1) * per, total)
        Wl = torch.zeros(per, W.shape[1], dtype=W.dtype, device=W.device)
        Wl[: hi - lo] = W[lo:hi]
        outs = ops.gptq_quantize(Wl, U, qt, stream=stream, **args)[:7]
        full = [torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in outs]
        # Stream of the gathers: the deferred group's side stream, else (GPU) a communication stream of this Quantizer, so
        # that the main stream can go on with the next group's column loop; _quant_group waits for it before the results
        # are consumed.
        gstream = stream
        if gstream is None and W.is_cuda:
            if self._gather_stream is None:
                self._gather_stream = torch.cuda.Stream()
            gstream = self._gather_stream
            self._gather_stream_used = True
        if gstream is not None:
            # Wl and the per-rank outputs die when this function returns, long before the other stream has run the kernel /
            # the all-gathers that use them: tell the caching allocator (otherwise the main stream's next allocations
            # -- the pass-2 activations -- would reuse the memory underneath the pending kernels).
            for t in (Wl,) + tuple(outs):
                if t is not None:
                    t.record_stream(gstream)
            gstream.wait_stream(torch.cuda.current_stream(W.device))     # kernel (if on main) done; gather buffers were allocated on main
        with (torch.cuda.stream(gstream) if gstream is not None else contextlib.nullcontext()):
            with self.timer.span("allgather"):
                for g, t in zip(full, outs):
                    dist.all_gather_into_tensor(g, t.contiguous())
        return tuple(g[:total] for g in full)

    # -------------------------------------------------------------------------------------------
    def _deferred_tail_plan(self, block, layers, hook_state, batches, device):
        """Can pass 2 of this block be split at its last quantised layer (out = residual + last(x))?

        Pre-norm decoder blocks (Llama family) end with `hidden = residual + mlp(norm(hidden))` where the MLP's last
        operation is the last quantised layer (down_proj) and `residual` is the input of `post_attention_layernorm`.
        That structure is not assumed but CHECKED once, on the first calibration batch of the first block: the output
        of the plain `block(...)` call must equal `residual + last(x)` bit for bit, otherwise the generic path is kept."""
        if not (self.defer_last_layer and self.overlap_prepare and not self.cpu_offload_activations):
            return None
        last_name = hook_state.get("last")
        norm = getattr(block, "post_attention_layernorm", None)
        if last_name is None or not isinstance(norm, nn.Module) or not next(block.parameters()).is_cuda:
            return None
        if not all(len(a) == 1 and isinstance(a[0], torch.Tensor) for a, _ in batches):
            return None
        plan = {"last_name": last_name, "last": layers[last_name], "norm": norm}
        if self._split_ok is None:
            inp_args, inp_kwargs = batches[0]
            cap = {}
            h1 = norm.register_forward_pre_hook(lambda m, a: cap.__setitem__("res", a[0]))
            h2 = plan["last"].register_forward_pre_hook(lambda m, a: cap.__setitem__("din", a[0]))
            try:
                full = maybe_first_element(block(*to(inp_args, device=device), **to(inp_kwargs, device=device)))
            finally:
                h1.remove(); h2.remove()
            ok = "res" in cap and "din" in cap and full.shape == cap["res"].shape
            if ok:
                ok = bool(torch.equal(full, cap["res"] + plan["last"](cap["din"])))
            self._split_ok = ok
            self._log(f"deferred-tail split of pass 2 at {last_name}: {'enabled' if ok else 'not applicable'}")
        return plan if self._split_ok else None

    def _split_fronts(self, block, tail, batches, device):
        """Run the block up to (not including) its last quantised layer for every batch; returns [(residual, x)]."""
        cap = {}

        def grab_res(_, a):
            cap["res"] = a[0]

        def grab_din(_, a):
            cap["din"] = a[0]
            raise ForwardInterrupt

        h1 = tail["norm"].register_forward_pre_hook(grab_res)
        h2 = tail["last"].register_forward_pre_hook(grab_din)
        fronts = []
        try:
            with self.timer.span("forward2"):
                for inp_args, inp_kwargs in batches:
                    try:
                        block(*to(inp_args, device=device), **to(inp_kwargs, device=device))
                    except ForwardInterrupt:
                        pass
                    fronts.append((cap.pop("res"), cap.pop("din")))
        finally:
            h1.remove(); h2.remove()
        return fronts

    # -------------------------------------------------------------------------------------------
    @staticmethod
    def _same_kwargs(k0, k) -> bool:
        if k.keys() != k0.keys():
            return False
        for key, v in k.items():
            v0 = k0[key]
            if isinstance(v, torch.Tensor):
                if not (isinstance(v0, torch.Tensor) and v.shape == v0.shape and torch.equal(v, v0)):
                    return False
            elif isinstance(v, (tuple, list)) and all(isinstance(t, torch.Tensor) for t in v):
                if not (isinstance(v0, (tuple, list)) and len(v0) == len(v) and
                        all(a.shape == b.shape and torch.equal(a, b) for a, b in zip(v, v0))):
                    return False
            elif v is not v0 and v != v0:
                return False
        return True

    def _batched_inputs(self, input_args, input_kwargs):
        """Merge per-sequence captures into batches: RUNS of consecutive sequences that only differ in hidden_states (same
        shape, equal keyword arguments) become one batch of up to `calibration_batch_size` sequences; anything else stays a
        batch of its own (fineweb_edu keeps the short tails of its documents, data_utils.py, so lengths may differ)."""
        bs = self.calibration_batch_size
        n = len(input_args)
        simple = bs > 1 and n > 1 and all(len(a) == 1 and isinstance(a[0], torch.Tensor) and a[0].shape[0] == 1 for a in input_args)
        if not simple:
            return [(list(a), k) for a, k in zip(input_args, input_kwargs)]
        out, i = [], 0
        while i < n:
            j = i + 1
            while (j < n and j - i < bs and input_args[j][0].shape == input_args[i][0].shape
                   and input_args[j][0].device == input_args[i][0].device and self._same_kwargs(input_kwargs[i], input_kwargs[j])):
                j += 1
            if j - i == 1:
                out.append((list(input_args[i]), input_kwargs[i]))
            else:
                out.append(([torch.cat([a[0] for a in input_args[i:j]], dim=0)], input_kwargs[i]))
            i = j
        return out

    def _front_capacity(self, tail, batches, device) -> int:
        """How many batches' (residual, last-layer input) pairs the deferred tail of pass 2 may keep on the GPU: they cost
        tokens x (hidden + in_features of the last layer) elements -- 120 GB for Llama-3-8B at 4M calibration tokens -- so the
        split is limited to half of the memory that is free right now; the remaining batches take the plain path after the
        deferred layer has finished."""
        try:
            free, _ = torch.cuda.mem_get_info(device)
        except Exception:  # noqa: BLE001
            return len(batches)
        budget = free // 2
        d_in = getattr(tail["last"], "in_features", 0)
        used, k = 0, 0
        for a, _ in batches:
            x = a[0]
            need = (x.numel() // x.shape[-1]) * (x.shape[-1] + d_in) * x.element_size()
            if used + need > budget:
                break
            used += need
            k += 1
        return k

    # -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def quantize(self, quant_config: Dict[str, GGMLQuantizationType]) -> None:
        """quantizer.py:59-217.  With `fused_forward_ops` the element-wise pieces of HF's block forward (RMSNorm, rotary
        embedding, SiLU*up) are swapped for single-pass kernels for the duration of the run (fused_forward.py: each
        replacement is verified against the module it replaces when installed, everything is restored afterwards)."""
        from .fused_forward import fused_forward
        ctx = fused_forward(self.model, self.verbose) if self.fused_forward_ops else contextlib.nullcontext([])
        with ctx as installed:
            self.fused_installed = list(installed)
            self._quantize_impl(quant_config)

    def _quantize_impl(self, quant_config: Dict[str, GGMLQuantizationType]) -> None:
        device = self.device or next(self.model.parameters()).device
        self._not_pd_flags = []
        self._mask_flags = []
        ops.set_timer(self.timer if self.timer.enabled else None)
        self._emit_counter = 0
        if self.save_dir is not None and (_rank() == 0 or self.spread_emission):
            os.makedirs(self.save_dir, exist_ok=True)
            self._saver = _AsyncSaver()
        blocks = self._get_submodule(self.block_modules)
        pre_blocks = [(n, self._get_submodule(n)) for n in self.pre_block_modules]
        post_blocks = [(n, self._get_submodule(n)) for n in self.post_block_modules]
        blocks[0] = blocks[0].to(device)
        for _, module in pre_blocks:
            module.to(device)
        use_cache = None
        if hasattr(self.model.config, "use_cache"):
            use_cache = self.model.config.use_cache
            self.model.config.use_cache = False

        # ---- capture the inputs of block 0 (quantizer.py:78-89) ----
        with self.timer.span("capture"):
            blocks[0] = InputCollector(blocks[0], cpu_offload=self.cpu_offload_activations)
            for inp_args, inp_kwargs in self.data_loader:
                try:
                    self.model(*to(inp_args, device=device), **to(inp_kwargs, device=device))
                except ForwardInterrupt:
                    pass
            input_args = blocks[0].input_args
            input_kwargs = blocks[0].input_kwargs
            blocks[0] = blocks[0].module
            batches = self._batched_inputs(input_args, input_kwargs)
            del input_args, input_kwargs
        if _dist_on():
            dist.barrier()

        # ---- pre-block modules (quantizer.py:94-128) ----
        for name, module in pre_blocks:
            if not self.quant_non_block_modules:
                continue
            self._log(f"Processing {name}.")
            self._quant_non_block_module(name, module.to(device), quant_config)
        if self.cpu_offload_modules:
            for _, module in pre_blocks:
                module.cpu()

        # ---- transformer blocks (quantizer.py:137-179) ----
        for block_id, block in enumerate(blocks):
            self._log(f"Processing {self.block_modules} {block_id}/{len(blocks)}.")
            block = block.to(device)
            layer_prefix = f"{self.block_modules}.{block_id}."
            layers = select_layers(self.model, layer_prefix, self.quantizable_modules, LINEAR_LAYERS)
            handles, hooks, end_of_forward, hook_state = self._prepare_hooks_and_handles(layers, quant_config, len(batches))

            with self.timer.span("forward1"):
                for inp_args, inp_kwargs in batches:
                    try:
                        block(*to(inp_args, device=device), **to(inp_kwargs, device=device))
                    except ForwardInterrupt:      # raised by the last hooked layer's pre-hook (pass-1 early exit)
                        pass
                    end_of_forward()
            for h in hooks.values():
                h.remove()

            tail = self._deferred_tail_plan(block, layers, hook_state, batches, device)
            finish = self._quant_group(handles, quant_config, defer=tail["last_name"] if tail else None,
                                       early=hook_state.get("early"))

            if finish is None:
                with self.timer.span("forward2"):
                    for inp_args, inp_kwargs in batches:
                        out = block(*to(inp_args, device=device), **to(inp_kwargs, device=device))
                        out = maybe_first_element(out)
                        if self.cpu_offload_activations:
                            out = out.cpu()
                        if len(inp_args) > 0:                                     # quantizer.py:167-168
                            if inp_args[0].shape == out.shape and inp_args[0].device == out.device:
                                inp_args[0].copy_(out)      # in place: batches are views of one activation buffer
                            else:
                                inp_args[0] = out
                        elif "hidden_states" in inp_kwargs:
                            inp_kwargs["hidden_states"] = out
                        else:
                            raise ValueError("Unsupported block input format.")
            else:
                # Pass 2, split at the block's last quantised layer: everything BEFORE it runs for all calibration
                # batches while that layer's Cholesky chain and column loop are still in flight on a side stream
                # (the block forwards are streams of short kernels, so the latency-bound chain interleaves with
                # them instead of idling the GPU); then the layer itself + the residual add for all batches.
                k_front = self._front_capacity(tail, batches, device)
                fronts = self._split_fronts(block, tail, batches[:k_front], device)
                finish()
                with self.timer.span("forward2"):
                    for (inp_args, _), (res, din) in zip(batches[:k_front], fronts):
                        out = res + tail["last"](din)
                        inp_args[0].copy_(out)
                    del fronts
                    for inp_args, inp_kwargs in batches[k_front:]:      # did not fit the memory budget: plain forward
                        out = maybe_first_element(block(*to(inp_args, device=device), **to(inp_kwargs, device=device)))
                        inp_args[0].copy_(out)
            if self.cpu_offload_modules:
                block = block.cpu()
            del handles, hooks

        # ---- post-block modules (quantizer.py:181-214) ----
        for name, module in post_blocks:
            if not self.quant_non_block_modules:
                continue
            self._log(f"Processing {name}")
            self._quant_non_block_module(name, module.to(device), quant_config)

        if use_cache is not None:
            self.model.config.use_cache = use_cache
        if self._saver is not None:
            with self.timer.span("save_wait"):
                self._saver.close()
            self._saver = None
        # Deferred validity check of the shared-Hessian groups (one device flag per group, no host sync inside the run): members
        # of a group must have the SAME all-zero weight columns, otherwise the shared U is wrong for some of them
        # (gptq.py:308-313 builds U per layer).  Every rank evaluates the same flags (W is replicated), the verdict is
        # all-reduced anyway so that no rank can leave the others waiting in the barrier below.
        bad_groups = [names for names, flag in self._mask_flags if bool(flag.item())]
        if _dist_on():
            verdict = torch.tensor([1 if bad_groups else 0], device=device if torch.device(device).type == "cuda" else "cpu")
            dist.all_reduce(verdict, op=dist.ReduceOp.MAX)
            if bool(verdict.item()) and not bad_groups:
                bad_groups = [["<reported by another rank>"]]
        ops.set_timer(None)
        if bad_groups:
            raise RuntimeError(f"layers {bad_groups} share an input but have different all-zero weight columns: the results "
                               f"written to {self.save_dir!r} for these layers are INVALID; rerun with share_hessians=False")
        if _dist_on():
            dist.barrier()

    def non_invertible_modules(self) -> List[str]:
        """Modules whose Hessian was not positive definite (U fell back to identity, gptq.py:321-323). Synchronises."""
        out = []
        for names, flags in self._not_pd_flags:
            if any(bool(f.item()) for f in flags):
                out.extend(names)
        return out
