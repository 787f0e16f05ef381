"""CPU restatement ("oracle") of pancax's deep-energy / FE-residual loss + gradient path.

THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / CPU baseline.  ``pancax_b200`` never
imports this package; the product path fails loudly if its CUDA library is missing.

What it restates (reference = /root/reference, pancax 0.0.16; file:line cited per function):
  * parent-element tables and quadrature rules      -> oracle/fem.py
  * Field / MLP forward                             -> oracle/field.py
  * constitutive energies (NeoHookean, Gent, Swanson, Hencky) and the
    sym-3x3 eigen / log-sqrt helpers                -> oracle/models.py
  * element interpolants, potential energy, internal force, residual norm,
    reaction force                                  -> oracle/physics.py
  * EnergyLoss, EnergyAndResidualLoss, EnergyResidualAndReactionLoss,
    FullFieldDataLoss and d(loss)/d(params)         -> oracle/losses.py

Arithmetic is float64 on the CPU: numpy for the tables, torch (CPU, float64) for
everything that must be differentiated -- torch.autograd plays the role jax.grad /
jax.jacfwd play in the reference, so gradients are "AD of the restated forward",
exactly how the reference obtains them.

Pinning status
--------------
The reference cannot be imported here (jax / equinox / optax are not installed and
there is no network) and its own test-suite never checks the loss/gradient path
numerically.  Therefore:
  * per component the oracle IS pinned: the reference's own pure-Python source for
    the parent tables, quadrature rules, constitutive energies, the sym-3x3 eigen
    solver / log-sqrt and ``potential_energy`` is executed in this container through
    a small numpy stand-in for the ``jax``/``equinox`` namespaces
    (``oracle/refshim``; generator ``oracle/gen_golden.py``; vectors in
    ``tests/golden/``), and the closed-form curves of the reference's
    ``test/constitutive_models`` tests are replayed as known-answer tests;
  * end to end (loss, aux, d loss / d theta, d loss / d prop_val) it is pinned too:
    ``oracle/refshim_torch.py`` backs the same namespaces with torch and maps every
    ``jax.grad`` / ``value_and_grad`` on the path to ``torch.autograd.grad``, so the
    reference's OWN loss functions, physics kernels, Field, function space and
    constitutive models are executed end to end (``oracle/gen_golden_ad.py`` ->
    ``tests/golden/loss_grad.npz``) and the oracle matches them to 1e-10
    (``tests/test_loss_grad_golden.py``).  Still restated, not executed:
    ``equinox.nn.MLP``/``Linear`` and optax's Adam (un-vendored third-party); not pinned
    against a running jax (none exists here).
"""
