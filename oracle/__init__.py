"""CPU oracle for the ELBO-gradient hot path of AdvancedVI.jl (v0.7.0 @ d3822cf).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker / the timed CPU
baseline.  The product path (``advancedvi.jl_b200``) never imports this package and
fails loudly when its CUDA library is missing.

What it is: a numpy restatement (float64 by default, float32 on request) of the
reference's arithmetic for the path SURVEY.md section 8(a) lists.  Every function cites
the reference file:line it follows (paths relative to /root/reference).

Pinning status ("parity partially pinned"):
  * The reference is 100 % Julia and Julia is not installed in this image, so the
    reference itself cannot be executed here and no outputs of it can be generated.
  * The oracle is pinned against every *known-answer* test the reference holds for
    this path (tests/test_oracle_known_answers.py re-expresses them: ELBO ~ 0 at q = pi,
    STL gradient ~ 0 at q = pi, epoch-mean minibatch gradient == full gradient,
    averaging weights, ClipScale bound, prox stationarity, family logpdf / entropy /
    moments against an independent scipy MvNormal, rule convergence), against
    closed-form expectations, against central finite differences of the restated
    forward passes, and (Philox) against the Random123 known-answer vectors.
  * It is additionally pinned against an INDEPENDENT AD backend and independent
    library arithmetic (tests/test_oracle_independent_ad.py): the reference's forward
    closures, its family and its documented regression models restated in PyTorch from
    the reference's source text and differentiated with autograd (values and gradients
    to 1e-10), torch.distributions for the densities, torch.optim for Adam / Descent.
  * NOT pinned by reference outputs (no golden vector exists in the reference): the RNG
    stream values (Julia Xoshiro/StableRNG + ziggurat randn is replaced by counter-based
    Philox4x32-10 + Box-Muller on both oracle and GPU) and the full-rank flatten layout;
    the docs-only logistic-regression arithmetic and the Optimisers.jl Adam arithmetic
    only through the independent implementations of the previous item.  DESIGN.md
    repeats this.
"""

from . import philox, family, models, objectives, optim, reshuffling  # noqa: F401
