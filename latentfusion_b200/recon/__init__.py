from . import models, fusion, inference, utils  # noqa: F401
from .models import Sculptor, Photographer  # noqa: F401
