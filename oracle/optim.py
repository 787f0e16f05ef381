"""Oracle (test infrastructure): the Adam update pancax composes in pancax/optimizers.py:155-173,
optax.chain(scale_by_adam(), scale_by_schedule(exponential_decay(lr, transition_steps, decay_rate)),
scale(-1.0)) -- optax is not vendored in /root/reference (unpinned dependency, pyproject.toml:27), so
its published algorithm is restated: scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0) with bias
correction by the incremented count; the schedule is evaluated at the pre-increment count."""
import numpy as np


class AdamState:
    def __init__(self, n):
        self.mu, self.nu, self.count = np.zeros(n), np.zeros(n), 0


def adam_step(p, g, st, lr0, transition_steps=500, decay_rate=0.99, b1=0.9, b2=0.999, eps=1e-8):
    lr = lr0 * decay_rate ** (st.count / transition_steps)
    st.mu = b1 * st.mu + (1 - b1) * g
    st.nu = b2 * st.nu + (1 - b2) * g * g
    st.count += 1
    mhat = st.mu / (1 - b1 ** st.count)
    vhat = st.nu / (1 - b2 ** st.count)
    return p - lr * mhat / (np.sqrt(vhat) + eps)
