"""Plugin base class and loader of the dloc toolbox convention (reference dloc/core/utils/base_model.py:8-46):
a model module holds exactly one BaseModel subclass; `conf` is merged over the class `default_conf`.

Inside the reference tree (its `dloc` package importable) the HOST's BaseModel / dynamic_load are re-exported, so that
plugins of this package are subclasses of the class the host's `dynamic_load` looks for; the definitions below are the
stand-alone fall-back (they have to restate the protocol: it is the interface)."""
try:
    from dloc.core.utils.base_model import BaseModel, dynamic_load  # noqa: F401  (host tree)
except ImportError:
    import inspect
    from abc import ABCMeta, abstractmethod
    from copy import copy

    from torch import nn

    class BaseModel(nn.Module, metaclass=ABCMeta):
        default_conf = {}
        required_data_keys = []

        def __init__(self, conf, model_path):
            super().__init__()
            self.conf = conf = {**self.default_conf, **conf}
            self.required_data_keys = copy(self.required_data_keys)
            self._init(conf, model_path)
            self.model_path = model_path

        def forward(self, data):
            for key in self.required_data_keys:
                assert key in data, 'Missing key {} in data'.format(key)
            return self._forward(data)

        @abstractmethod
        def _init(self, conf, model_path):
            raise NotImplementedError

        @abstractmethod
        def _forward(self, data):
            raise NotImplementedError

    def dynamic_load(root, model):
        """Import `<root>.<model>` and return its single BaseModel subclass."""
        module_path = f'{root.__name__}.{model}'
        module = __import__(module_path, fromlist=[''])
        found = [cls for _, cls in inspect.getmembers(module, inspect.isclass)
                 if cls.__module__ == module_path and issubclass(cls, BaseModel)]
        assert len(found) == 1, found
        return found[0]
