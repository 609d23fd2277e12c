def ddeint_adjoint():
    raise NotImplementedError  # paddlexde/functional/ddeint_adjoint.py:1-2
