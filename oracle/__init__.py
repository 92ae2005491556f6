"""CPU oracle for the Act3D / ChainedDiffuser hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as
the checker / CPU baseline, never as the thing shipped or measured as ours.

What it is: a functional, plain-PyTorch-on-CPU restatement (fp32 by default,
fp64 on request) of the reference algorithm for the path named by
BASELINE.json's north_star, written against ``state_dict`` tensors so that it
can be driven by either the reference's modules or ours (same keys).  Every
function cites the reference file:line it follows (paths relative to
/root/reference).

How it is pinned: ``tests/golden/make_golden.py`` imports the *unmodified*
reference from /root/reference in the build container (with import stubs for
the two packages that are not installed, ``clip`` and ``diffusers``), runs it
on seeded synthetic inputs, and commits the outputs under ``tests/golden``.
``tests/test_oracle_golden.py`` checks this oracle against those vectors.

Parity status
  * Act3D path, attention stacks, RoPE, top-k, diffusion head: pinned by the
    golden vectors above (generated from the reference itself).
  * ``DDPMScheduler`` (third-party ``diffusers``, unpinned in the reference's
    README.md:29 and not installed here): restated from its published
    algorithm in ``oracle/ddpm.py`` -- **parity unpinned** for that component;
    it is anchored by closed-form known answers only (tests/test_ddpm.py).
  * OpenAI ``clip`` RN50 backbone (README.md:30): not installed, weights not
    available offline -- **parity unpinned**; the oracle and the product use
    ``backbone="resnet"`` (torchvision, identical library code on both sides).
"""
