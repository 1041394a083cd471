"""CPU-only helpers (no device calls): circuits as dense matrices for the oracle."""
from helpers import circuit_for_oracle, tfim_layer


def tfim_layer_cpu(g, dt=0.25, hx=1.0, hz=0.8, J=0.5, ncol=4):
    layer = tfim_layer(g, dt, hx, hz, J, ncol)
    gm, gv = circuit_for_oracle(g, layer)
    return layer, gm, gv
