"""forte2_b200 — B200-native determinant-CI sigma builder behind forte2's CIStrings /
CISigmaBuilder interface.  See DESIGN.md and INTEGRATION.md."""
from .ci_strings import CIStrings  # noqa: F401
from .ci_sigma_builder import CISigmaBuilder  # noqa: F401
from .ci_spin_adapter import CISpinAdapter  # noqa: F401

__all__ = ["CIStrings", "CISigmaBuilder", "CISpinAdapter"]
