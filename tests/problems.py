"""Shared synthetic problems (BASELINE.json configs) -- NumPy only."""
import numpy as np


def spiral_weights(d=2, h=50, seed=42, scale=0.1):
    """ODEFunc init (example/ode_demo.py:27-30): W ~ 0.1*N(0,1), b = 0; default_rng(seed) stands in
    for paddle.seed(42)."""
    rng = np.random.default_rng(seed)
    w1 = (scale * rng.standard_normal((d, h))).astype(np.float32)
    w2 = (scale * rng.standard_normal((h, d))).astype(np.float32)
    return w1, np.zeros(h, np.float32), w2, np.zeros(d, np.float32)


def fanin_weights(d, h, seed=1):
    """cfg3/cfg4: W ~ N(0,1)/sqrt(fan_in), small random biases."""
    rng = np.random.default_rng(seed)
    w1 = (rng.standard_normal((d, h)) / np.sqrt(d)).astype(np.float32)
    b1 = (0.1 * rng.standard_normal(h)).astype(np.float32)
    w2 = (rng.standard_normal((h, d)) / np.sqrt(h)).astype(np.float32)
    b2 = (0.1 * rng.standard_normal(d)).astype(np.float32)
    return w1, b1, w2, b2


def cfg2_y0(B, seed=0):
    """cfg2: y0 = [2,0] + 0.5*N(0,1)"""
    rng = np.random.default_rng(seed)
    return (np.array([2.0, 0.0]) + 0.5 * rng.standard_normal((B, 2))).astype(np.float32)


def cfg2_tspan(n=10):
    return np.linspace(0.0, 25.0, 1000).astype(np.float32)[:n]


def spiral_truth(n=1000):
    """SimpleDemoData (example/demo_utils.py:147-164): dy/dt = y^3 A, y0=[2,0], RK4 on linspace(0,25,n)."""
    A = np.array([[-0.1, 2.0], [-2.0, -0.1]], np.float32)
    t = np.linspace(0.0, 25.0, n).astype(np.float32)
    return t, A


def dde_field_coefficients():
    """The stand-in delay field of the reference-run ddeint vectors: dy = 0.5 * y_lags - 0.25 * y, elementwise fp32 (two
    rounded products and one rounded difference: the same bits from NumPy, the paddle stand-in and torch eager)."""
    return np.float32(0.5), np.float32(0.25)
