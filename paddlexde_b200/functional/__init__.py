from .odeint import odeint  # noqa: F401
from .odeint_adjoint import odeint_adjoint  # noqa: F401
from .sdeint import sdeint  # noqa: F401
from .ddeint import ddeint  # noqa: F401
from .ddeint_adjoint import ddeint_adjoint  # noqa: F401
from .sdeint_adjoint import sdeint_adjoint  # noqa: F401
