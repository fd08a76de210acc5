"""Common base of the model classes (role of models/base.py:1-31): abstract forward, trainable-parameter count in str()."""
import torch.nn as nn


class BaseModel(nn.Module):
    # run-time state of event_flow_b200.graphed (CUDA graphs of the no-grad step): never part of a checkpoint or a deepcopy
    _GRAPH_KEYS = ("_step_graphs", "_graph_params", "_graph_off", "_graph_error", "_no_states")

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in self._GRAPH_KEYS:
            state.pop(k, None)
        return state

    def forward(self, *inputs):
        raise NotImplementedError(f"{type(self).__name__} does not define forward()")

    def trainable_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def __str__(self):
        return f"{super().__str__()}\nTrainable parameters: {self.trainable_parameters()}"
