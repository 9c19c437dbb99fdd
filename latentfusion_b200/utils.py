"""Small host utilities: block-config DSL parser and loss-weight schedulers.
API mirror of the parts of reference ``latentfusion/utils.py`` the path consumes
(parse_block_config :40-50, Exponential/LinearScheduler)."""
import math


def parse_block_str(s):
    return s if s in {'I', 'U', 'D'} else int(s)


def parse_block_config(s, delimiter=',', group_delimiter=':'):
    """"64,D,128:128,U,64" -> [[64,'D',128],[128,'U',64]];  "32,32" -> [32,32];  ""/"none" -> []."""
    if s.lower() == 'none' or len(s) == 0:
        return []

    def parse(section):
        return [parse_block_str(tok) for tok in section.split(delimiter)] if section else []
    if group_delimiter in s:
        return [parse(section) for section in s.split(group_delimiter)]
    return parse(s)


class LinearScheduler:
    """value(step) = lerp(initial, end, step/num_steps) (not clamped, like the reference)."""

    def __init__(self, initial_value, end_value, num_steps):
        self.initial_value, self.end_value, self.num_steps = initial_value, end_value, num_steps

    def get(self, step):
        a = step / self.num_steps
        return (1.0 - a) * self.initial_value + a * self.end_value


class ExponentialScheduler:
    """Exponential decay hitting final_value at step num_steps-1, constant afterwards."""

    def __init__(self, initial_value, final_value, num_steps):
        self.initial_value, self.final_value, self.num_steps = initial_value, final_value, num_steps
        self.mean_lifetime = -(num_steps - 1) / math.log(final_value / initial_value)

    def get(self, step):
        if step >= self.num_steps:
            return self.final_value
        return self.initial_value * math.exp(-step / self.mean_lifetime)


def trange(n, **kwargs):
    try:
        from tqdm.auto import trange as _trange
        return _trange(n, **kwargs)
    except Exception:      # pragma: no cover
        return range(n)
