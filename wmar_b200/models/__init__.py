from .armm_wrapper import AutoregressiveMultimodalModelWrapper  # noqa: F401
from .rar_wrapper import RarARMMWrapper  # noqa: F401
from .taming_wrapper import TamingARMMWrapper  # noqa: F401
from .chameleon_wrapper import ChameleonARMMWrapper  # noqa: F401
