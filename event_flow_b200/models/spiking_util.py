"""
Spike-function names of models/spiking_util.py:96-109.  In this package the Heaviside forward and the surrogate
backward live inside the fused CUDA kernels (csrc/common.cuh: neuron_update, surrogate_grad); the callables below are
only markers the cells resolve with getattr(spiking, activation), exactly like the reference does.
"""


class _SpikeMarker:
    def __init__(self, name, default_width):
        self.name, self.default_width = name, default_width

    def __call__(self, *a, **k):
        raise RuntimeError(
            f"{self.name} is fused into the conv+neuron CUDA kernel in event_flow_b200 and cannot be called stand-alone"
        )

    def __repr__(self):
        return f"<fused spike fn {self.name}>"


superspike = _SpikeMarker("superspike", 10.0)
mgspike = _SpikeMarker("mgspike", 0.5)
trianglespike = _SpikeMarker("trianglespike", 1.0)
arctanspike = _SpikeMarker("arctanspike", 10.0)
