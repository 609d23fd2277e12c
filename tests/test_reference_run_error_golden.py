"""The reference's assertions on the adaptive path (solver/base_adaptive_solver_rk.py:120-122, 200-203), produced by its own
unmodified code on the NumPy `paddle` stand-in (tools/make_reference_error_golden.py): `max_num_steps exceeded`,
`non-finite values in state`, `underflow in dt` -- incl. a field that returns NaN from a finite state (dt becomes NaN) and
a controller that halves dt down to 0.  The oracle reports the matching status word after the same number of completed
attempts; the package's shim turns the status word back into an AssertionError that starts like the reference's."""
import ast
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Z = np.load(os.path.join(ROOT, "tests", "golden", "reference_run_errors.npz"), allow_pickle=False)
NAMES = sorted({k.split("/")[0] for k in Z.files})
EXPECT = {"max_num_steps exceeded": 3, "non-finite values in state `y`": 2, "underflow in dt": 1}  # XDE_ST_* / ORC_*


def expected(name):
    msg = str(Z[f"{name}/message"])
    (prefix, code), = [(p, c) for p, c in EXPECT.items() if msg.startswith(p)]
    # the reference counts the call of _adaptive_step that raised; max_num_steps is asserted before the call (:120-122)
    done = int(Z[f"{name}/adaptive_step_calls"]) - (0 if code == 3 else 1)
    return prefix, code, done


@pytest.mark.parametrize("name", NAMES)
def test_oracle_status_words_follow_the_reference_assertions(oracle, name):
    from paddlexde_b200._lib import raise_for_status

    meta = ast.literal_eval(str(Z[f"{name}/meta"]))
    om = oracle.MLP(Z[f"{name}/w1"], Z[f"{name}/b1"], Z[f"{name}/w2"], Z[f"{name}/b2"], pre=meta.pop("pre"))
    prefix, code, done = expected(name)
    out, st, _, rc = oracle.dopri5_mlp(om, Z[f"{name}/y0"], Z[f"{name}/t"], controller="batch", **meta)
    assert rc == code and int(st.n_attempts[0]) == done
    rows = Z[f"{name}/rows_done"]
    assert np.array_equal(out[:len(rows)], rows, equal_nan=True)
    with pytest.raises(AssertionError) as e:
        raise_for_status(rc)
    assert str(e.value).startswith(prefix)  # the shim's text (paddlexde_b200/_lib.py) starts like the reference's


def test_committed_error_vectors_are_what_the_reference_does_here():
    from oracle.ref_shim import loader

    if not loader.available():
        pytest.skip("/root/reference is not on this machine")
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_reference_error_golden",
                                                  os.path.join(ROOT, "tools", "make_reference_error_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # the NaN cases divide by NaN on purpose
        out, _ = gen.generate()
    assert sorted(out) == sorted(Z.files)
    for k in Z.files:
        a, b = np.asarray(out[k]), Z[k]
        assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes(), k
