"""Common base of the model classes (role of models/base.py:1-31): abstract forward, trainable-parameter count in str()."""
import torch.nn as nn


class BaseModel(nn.Module):
    def forward(self, *inputs):
        raise NotImplementedError(f"{type(self).__name__} does not define forward()")

    def trainable_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def __str__(self):
        return f"{super().__str__()}\nTrainable parameters: {self.trainable_parameters()}"
