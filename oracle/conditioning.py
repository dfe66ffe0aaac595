"""Test infrastructure (NOT a product path): well-conditioned weights for gradient parity tests.

The gradient of a ReLU network is discontinuous wherever a pre-activation is exactly zero: two
correct implementations whose forward values differ by one rounding error disagree on the gate of
any unit whose pre-activation lies inside that error, and each disagreement moves whole gradient
entries by O(1/sqrt(rows)) of the tensor's scale.  This is a property of the function, not of an
implementation: on the C1 configuration at random initial weights this oracle evaluated in float32
differs from itself in float64 by 1.7e-3 (max-norm, `enc.1.w1`) through a single flipped gate
(tests/test_oracle.py::test_relu_gate_flips_limit_float32_agreement).

Parity at the 1e-3 level is therefore measured at weights for which no ReLU pre-activation on the
test batch - encoder feed-forward units (transformer.py:163-167) and head MLP units
(head.py:16-19, :41-43) - lies within `tau` of zero.  `condition_relu_gates` gets there by
nudging the biases of the offending units (float32-representable values, ~1e-2 in total at most),
re-running the float64 forward until the batch is clear.
"""
import numpy as np

from . import clickpath_oracle as O


def relu_margins(P, ids_list, num_layers, num_heads, pe, head_rows):
    """[(bias name, |pre-activation| array (rows, units) over the rows that matter)] for every
    ReLU stage, in forward order.  head_rows(x, ids_first) -> (n, d) rows fed to the head MLP."""
    x, caches = O.encoder_fwd(ids_list, P, num_layers, num_heads, pe, np.float64)
    live = (ids_list[0] != O.INPUT_PAD).reshape(-1)     # pad rows are dead compute
    out = []
    for l, (c, _) in enumerate(caches):
        pre = c["pre"].reshape(-1, c["pre"].shape[-1])
        out.append((f"enc.{l}.b1", np.abs(pre[live])))
    rows = head_rows(x, ids_list[0])
    i = 0
    while f"head.{i}.w" in P:
        pre = rows @ P[f"head.{i}.w"] + P[f"head.{i}.b"]
        out.append((f"head.{i}.b", np.abs(pre)))
        rows = np.maximum(pre, 0)
        i += 1
    return out


def masked_rows(x, ids_first):
    sel, _ = O.select_masked(ids_first, x)
    return sel.reshape(-1, x.shape[-1])


def segment_rows(segment):
    def f(x, ids_first):
        starts, ends = O.segment_bounds(ids_first[0])
        return x[:, int(starts[segment]):int(ends[segment]), :].reshape(-1, x.shape[-1])
    return f


def condition_relu_gates(P, ids_list, num_layers, num_heads, pe, head_rows=masked_rows, tau=1e-4,
                         seed=0, max_iter=200):
    """Copy of P (float64 arrays holding float32-representable values) in which the biases of
    ReLU units with a pre-activation inside (-tau, tau) on this batch have been nudged until no
    such unit is left.  Returns (P, number of nudged units)."""
    rng = np.random.default_rng(seed)
    P = {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in P.items()}
    nudged = 0
    for _ in range(max_iter):
        stages = relu_margins(P, ids_list, num_layers, num_heads, pe, head_rows)
        bad = [(name, np.flatnonzero((a < tau).any(axis=0))) for name, a in stages]
        bad = [(n, u) for n, u in bad if len(u)]
        if not bad:
            return P, nudged
        name, units = bad[0]            # earliest stage first: later stages see its change
        b = P[name].copy()
        b[units] += rng.choice([-1.0, 1.0], size=len(units)) * rng.uniform(10 * tau, 40 * tau, size=len(units))
        P[name] = b.astype(np.float32).astype(np.float64)
        nudged += len(units)
    raise RuntimeError("condition_relu_gates did not converge")


def min_relu_margin(P, ids_list, num_layers, num_heads, pe, head_rows=masked_rows):
    return min(float(a.min()) for _, a in relu_margins(P, ids_list, num_layers, num_heads, pe, head_rows)
               if a.size)
