"""Whole-step CUDA graph for the train loop (zero_grad → forward(loss) → backward → fused SGD step).

The step of /root/reference/model/bisenet/cityscapes.bisenet.R18/train.py:116-142 is ~400 kernel launches of 5-200 us
each; captured once, a replay removes the per-launch CPU work (Python autograd, ctypes, tensor-map encoding) and the
inter-kernel launch gaps. Single process / single GPU only (the DDP side-stream all-reduce is not captured).

Usage:
    step = GraphedTrainStep(model, optimizer, example_inputs)     # runs `warmup` eager steps, then captures
    for it, batch in enumerate(loader):
        set the learning rates in optimizer.param_groups as usual
        loss = step(batch['data'], batch['label'])                # device tensors; copied into the static inputs
If capture is impossible (non-CUDA device, an op that cannot be captured) the object runs eager steps and keeps the
reason in `.error`.

Status: EXPERIMENTAL. Verified on B200: capture + replay follow the eager trajectory (tests/test_gpu_bisenet.py::
test_graphed_train_step_matches_eager; tools/diag_graph*.py at the bench size: 329 launches per graph), but in the full
bench.py flow the capture was intermittently invalidated (cudaErrorStreamCaptureInvalidated) by a call not yet
identified, and a failed capture leaves the eager step ~50 % slower and the process-wide CUDA RNG in capture mode — so
bench.py keeps eager launches by default (`--graph` opts in) and the GPU test only runs with TSB_TEST_GRAPH=1.
What the B200 runs showed (tools/diag_graph*.py, tests): captures preceded by warm-up steps on a SIDE stream succeeded
(first version of the test, diag_graph2 VAR b), captures preceded only by default-stream steps failed (bench flow, the
test with a hand-written default-stream warm-up) — consistent with PyTorch's documented requirement that the warm-up
run on a side stream (autograd's leaf-gradient streams must not be the legacy default stream, whose implicit use inside
a global-mode capture invalidates it); the side-stream warm-up is therefore the default again, with `warmup >= 1`
required. Ruled out on hardware: cyclic GC during capture (now collected before / disabled during the capture anyway)
and the pinned-host → device hyper-parameter copy inside the graph (removed: hyper-parameters live in a device tensor
refreshed by an eager copy before each replay)."""
import torch


class GraphedTrainStep(object):
    def __init__(self, model, optimizer, example_inputs, warmup=3, enable=True, side_stream_warmup=True,
                 capture_error_mode="global"):
        self.model = model
        self.opt = optimizer
        self.static_inputs = [t.clone() for t in example_inputs]
        self.graph = None
        self.static_loss = None
        self.launches_per_step = 0
        self.error = None
        dev = self.static_inputs[0].device
        if not enable or dev.type != "cuda":
            return
        if side_stream_warmup and warmup < 1:
            warmup = 1       # at least one step must have run on a non-default stream before the capture
        from .. import _lib
        try:
            if side_stream_warmup:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(warmup):
                        self._eager(*self.static_inputs)
                torch.cuda.current_stream(dev).wait_stream(side)
            else:
                for _ in range(warmup):          # lazy attributes, FlatPack mirror, per-thread context binding all warm
                    self._eager(*self.static_inputs)
            torch.cuda.synchronize(dev)
            import gc
            self.opt.push_hyperparams()
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            # torch.cuda.graph no longer garbage-collects on entry: collect now and keep the cyclic GC off while capturing,
            # so that no CUDA object (tensor of an old autograd graph, stream, event) is destroyed in the middle of it
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g, capture_error_mode=capture_error_mode):
                    self.static_loss = self._eager(*self.static_inputs)
            finally:
                if gc_was_on:
                    gc.enable()
            self.launches_per_step = _lib.launch_count() - n0
            self.graph = g
        except Exception as e:  # noqa: BLE001 — fall back to eager steps, keep the reason
            self.error = "%s: %s" % (type(e).__name__, e)
            self.graph = None
            try:                 # best effort: release the half-built graph (its RNG registration, its memory pool)
                g.reset()
            except Exception:    # noqa: BLE001
                pass
            try:
                torch.cuda.synchronize(dev)
            except Exception:    # noqa: BLE001
                pass

    def _eager(self, *inputs):
        self.opt.zero_grad()
        loss = self.model(*inputs)
        loss.backward()
        self.opt.step()
        return loss

    def __call__(self, *inputs):
        if self.graph is None:
            return self._eager(*inputs)
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.opt.push_hyperparams()      # lr / weight decay of this step → the device tensor the captured SGD kernel reads
        self.graph.replay()
        # Python-side caches describe the weights BEFORE the replayed SGD update: drop them so that an eager step (or
        # evaluation) after a replay re-packs from the current parameters
        from .. import ops
        ops.pack_cache.cache.clear()
        ops.pack_cache.step += 1
        pack = getattr(self.opt, "pack", None)
        if pack is not None:
            pack.mark_fresh()
        return self.static_loss
