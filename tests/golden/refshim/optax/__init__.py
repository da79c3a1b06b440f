"""optax stand-in: the update rules of optax 0.1.2 / 0.1.5 as documented, float64.

adam(lr, b1=.9, b2=.999, eps=1e-8, eps_root=0):  mu = b1 mu + (1-b1) g;  nu = b2 nu + (1-b2) g^2;  count += 1;
    update = -lr * (mu / (1-b1^count)) / (sqrt(nu / (1-b2^count) + eps_root) + eps)
sgd(lr, momentum):  trace = g + momentum * trace;  update = -lr * trace     (nesterov=False)
adagrad(lr, initial_accumulator_value=0.1, eps=1e-7):  acc += g^2;  update = -lr * g / sqrt(acc + eps)
"""
import collections

import torch as _t

import jax as _jax

GradientTransformation = collections.namedtuple("GradientTransformation", "init update")
ScaleByAdamState = collections.namedtuple("ScaleByAdamState", "count mu nu")
TraceState = collections.namedtuple("TraceState", "trace")
ScaleByRssState = collections.namedtuple("ScaleByRssState", "sum_of_squares")
EmptyState = collections.namedtuple("EmptyState", "")

_map = _jax.tree_map


def apply_updates(params, updates):
    return _map(lambda p, u: p + u, params, updates)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(params):
        z = lambda p: _t.zeros_like(p)                             # noqa: E731
        return (ScaleByAdamState(count=0, mu=_map(z, params), nu=_map(z, params)), EmptyState())

    def update(grads, state, params=None):
        s = state[0]
        mu = _map(lambda g, m: b1 * m + (1 - b1) * g, grads, s.mu)
        nu = _map(lambda g, v: b2 * v + (1 - b2) * g * g, grads, s.nu)
        count = s.count + 1
        c1, c2 = 1 - b1 ** count, 1 - b2 ** count
        upd = _map(lambda m, v: -learning_rate * (m / c1) / (_t.sqrt(v / c2 + eps_root) + eps), mu, nu)
        return upd, (ScaleByAdamState(count=count, mu=mu, nu=nu), EmptyState())
    return GradientTransformation(init, update)


def sgd(learning_rate, momentum=None, nesterov=False):
    assert not nesterov

    def init(params):
        if momentum is None:
            return (EmptyState(), EmptyState())
        return (TraceState(trace=_map(lambda p: _t.zeros_like(p), params)), EmptyState())

    def update(grads, state, params=None):
        if momentum is None:
            return _map(lambda g: -learning_rate * g, grads), state
        tr = _map(lambda g, t: g + momentum * t, grads, state[0].trace)
        return _map(lambda t: -learning_rate * t, tr), (TraceState(trace=tr), EmptyState())
    return GradientTransformation(init, update)


def adagrad(learning_rate, initial_accumulator_value=0.1, eps=1e-7):
    def init(params):
        return (ScaleByRssState(sum_of_squares=_map(lambda p: _t.full_like(p, initial_accumulator_value), params)),
                EmptyState())

    def update(grads, state, params=None):
        acc = _map(lambda g, a: a + g * g, grads, state[0].sum_of_squares)
        upd = _map(lambda g, a: -learning_rate * g / _t.sqrt(a + eps), grads, acc)
        return upd, (ScaleByRssState(sum_of_squares=acc), EmptyState())
    return GradientTransformation(init, update)
