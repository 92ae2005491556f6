"""Whole-step CUDA-graph capture for the training loops (engine.py:95-160 / main_keypose.py:207-229 drive
``loss = criterion(model(...)); loss.backward(); optimizer.step()`` once per batch).

A training step of Act3D at the reference's batch sizes is ~1500 short kernels: the host needs as long to launch
them (~38 ms) as the GPU needs to run them.  The step has no host synchronisation (device-side ghost sampler,
top-k, argmax), fixed shapes, and a capturable optimizer, so forward + loss + backward + AdamW + (under DDP) the
bucketed NCCL all-reduce are captured once and replayed; a step then costs one graph launch plus the copy of the
batch into the static input buffers.

DDP (torch.nn.parallel.DistributedDataParallel) is captured with its all-reduce inside the graph; this needs
  * ``find_unused_parameters=False`` -- use :func:`freeze_parameters_without_gradient` first: the reference relies
    on find_unused_parameters=True only because six FPN output blocks never reach the loss;
  * the DDP wrapper constructed on a side stream (:func:`build_ddp`) and warmed up for >= 11 iterations before the
    capture (:class:`GraphedTrainStep` does the warm-up);
  * ``TORCH_NCCL_ASYNC_ERROR_HANDLING=0`` in the environment before ``init_process_group`` (the watchdog thread must
    not poll the captured collectives).
"""
import torch


def freeze_parameters_without_gradient(model, step_loss):
    """One eager forward/backward; every trainable parameter whose ``.grad`` stays None is set to
    ``requires_grad=False`` (it is unreachable from the loss).  Returns their names."""
    for p in model.parameters():
        p.grad = None
    step_loss(model).backward()
    frozen = []
    for name, p in model.named_parameters():
        if p.requires_grad and p.grad is None:
            p.requires_grad_(False)
            frozen.append(name)
        p.grad = None
    return frozen


def build_ddp(model, **kwargs):
    """DistributedDataParallel constructed on a side stream, as whole-step capture requires (its buckets and hooks must
    not belong to the stream the graph is later captured on)."""
    dev = next(model.parameters()).device
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        net = torch.nn.parallel.DistributedDataParallel(model, **kwargs)
    main.wait_stream(side)
    return net


class GraphedTrainStep:
    """step_loss(net, *inputs) -> scalar loss.  ``inputs``: example CUDA tensors (shapes / dtypes are fixed).
    Call the object with a new batch: copies it into the static buffers, replays the graph and returns the (static)
    loss tensor of that step.  ``make_optimizer(params)`` must build a capturable optimizer, e.g.
    ``torch.optim.AdamW(params, lr=..., capturable=True)``.  ``net``: the module to call -- ``model`` itself, or a
    DDP wrapper made by :func:`build_ddp` (then the bucketed all-reduce is part of the graph)."""

    def __init__(self, model, step_loss, make_optimizer, inputs, net=None, warmup=11):
        dev = inputs[0].device
        self.model = model
        self.net = model if net is None else net
        self.static_in = [t.detach().clone() for t in inputs]
        main = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self.optimizer = make_optimizer([p for p in model.parameters() if p.requires_grad])
            for _ in range(warmup):
                self.optimizer.zero_grad(set_to_none=True)
                step_loss(self.net, *self.static_in).backward()
                self.optimizer.step()
        main.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        if hasattr(model, "ensure_sampler_counter"):
            model.ensure_sampler_counter(dev)          # allocated outside the capture (its zero-fill must not be replayed)
        scope = model.device_sampler_counter(dev) if hasattr(model, "device_sampler_counter") else _null()
        with torch.cuda.graph(self.graph):
            with scope:
                self.loss = step_loss(self.net, *self.static_in)
            self.loss.backward()
            self.optimizer.step()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
