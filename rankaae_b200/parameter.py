"""Config namespace and class registries (mirrors `sc/utils/parameter.py:14-94` of the reference).

`Parameters` keeps the reference's behaviour exactly (immutable attributes, `.get`, `.update`,
`.to_dict`, `.from_yaml`; pinned by the reference's own sc/tests/test_parameters.py, restated in
tests/test_parameters.py).  The registries contain what the fused path implements: the FC family
and AdamW; asking for anything else raises at construction time instead of silently running a
different code path.
"""
from torch import optim

from .model import FCDecoder, FCEncoder

AE_CLS_DICT = {
    "FC": {"encoder": FCEncoder, "decoder": FCDecoder},
}

OPTIM_DICT = {
    "AdamW": optim.AdamW,
}


class Parameters():
    """A parameter object that maps all dictionary keys into its name space (namedtuple-like)."""

    def __init__(self, parameter_dict):
        super().__setattr__("_parameter_dict", parameter_dict)
        self.update(parameter_dict)

    def __setattr__(self, __name, __value):
        raise TypeError('Parameters object cannot be modified after instantiation')

    def get(self, key, value):
        return self._parameter_dict.get(key, value)

    def update(self, parameter_dict):
        self._parameter_dict.update(parameter_dict)
        self.__dict__.update(self._parameter_dict)

    def to_dict(self):
        return self._parameter_dict

    @classmethod
    def from_yaml(cls, config_file_path):
        import yaml

        with open(config_file_path) as f:
            trainer_config = yaml.full_load(f)
        return Parameters(trainer_config)
