"""CPU oracle for the gyre sampling hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch fp32 restatement of the arithmetic the reference
(stablecabal/gyre @ 9cba9781) executes on its diffusion sampling path: the
diffusers-0.16 ``UNet2DConditionModel`` / ``AutoencoderKL`` op graph, gyre's CFG
wrapper, the k-diffusion / DDIM scheduler inner loops and the ToMe K/V merge.

PARITY STATUS: "parity unpinned" at the diffusers boundary.  The reference holds no
golden tensors / known-answer tests for this path (every test in its ``tests/`` writes
PNGs for eyeballing) and ``diffusers`` itself is a third-party dependency that is
neither vendored in the reference tree nor installed in this image.  What IS pinned:

* the sampler loops and denoiser wrappers against the *vendored*
  ``gyre/src/k-diffusion/k_diffusion/{sampling,external}.py`` (imported from
  ``/root/reference`` by ``scripts/make_golden.py``; vectors in ``tests/golden``),
* the ToMe merge against the *vendored* ``nonfree/ToMe/tome/merge.py``,
* the DDIM step against the in-tree copy ``gyre/pipeline/schedulers/scheduling_ddim.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product (``gyre_b200``) never
does, and fails loudly if its CUDA library is missing.
"""
