"""gpytorch.module.Module: parameter / constraint registration and ``initialize``."""
import torch
from torch import nn


class Module(nn.Module):
    def __init__(self):
        super().__init__()
        self._added_loss_terms = {}
        self._priors = {}
        self._constraints = {}

    def forward(self, *inputs, **kwargs):
        raise NotImplementedError

    def __call__(self, *inputs, **kwargs):
        outputs = self.forward(*inputs, **kwargs)
        if isinstance(outputs, list):
            return [o for o in outputs]
        return outputs

    def register_parameter(self, name, parameter):
        if "_parameters" not in self.__dict__:
            raise AttributeError("Cannot assign parameter before Module.__init__() call")
        super().register_parameter(name, parameter)

    def register_constraint(self, param_name, constraint, replace=True):
        if param_name not in self._parameters:
            raise RuntimeError("Attempting to register constraint for nonexistent parameter.")
        constraint_name = param_name + "_constraint"
        self.add_module(constraint_name, constraint)
        self._constraints[constraint_name] = constraint
        if constraint.initial_value is not None:
            self.initialize(**{param_name: constraint.inverse_transform(constraint.initial_value)})

    def initialize(self, **kwargs):
        """gpytorch.Module.initialize: properties go through their setter, tensors are copied (expanded),
        floats fill."""
        for name, val in kwargs.items():
            if isinstance(val, int):
                val = float(val)
            if "." in name:
                module, name = name.rsplit(".", 1)
                self.get_submodule(module).initialize(**{name: val})
            elif not hasattr(self, name):
                raise AttributeError(f"Unknown parameter {name} for {self.__class__.__name__}")
            elif name not in self._parameters and name not in self._buffers:
                setattr(self, name, val)
            elif torch.is_tensor(val):
                self.__getattr__(name).data.copy_(val.expand_as(self.__getattr__(name)))
            elif isinstance(val, float):
                self.__getattr__(name).data.fill_(val)
            else:
                raise AttributeError(f"Type {type(val)} not valid for initializing parameter {name}")
        return self
