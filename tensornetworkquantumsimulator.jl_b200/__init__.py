"""tnqs-b200: B200-native belief-propagation simple-update engine (host mirror of the reference API).

Import as `tnqs_b200` (see /tnqs_b200.py at the repo root)."""
from . import graphs, gates  # noqa: F401
from .graphs import (NamedGraph, named_grid, named_path_graph, named_comb_tree, eagle_heavy_hex,  # noqa: F401
                     build_graph_from_gates, build_graph_from_circuit, edge_color,
                     forest_cover_edge_sequence, bipartite_edge_sequence)
from .gates import (ArgumentError, gate_matrix, observable_matrix, register_gate, register_alias,  # noqa: F401
                    unregister_gate)
from .api import (TensorNetworkState, BeliefPropagationCache, tensornetworkstate, zerostate,  # noqa: F401
                  random_tensornetworkstate, random_bpc_on_device, apply_gates, apply_circuit, truncate, update, expect, network,
                  maxvirtualdim, messages, message, default_bp_update_kwargs, circuit_arrays,
                  vertex_scalar, vertex_scalars, edge_scalar, edge_scalars, scalar_factors_quotient, freenergy,
                  partitionfunction, rescale_messages, rescale_vertices, rescale, norm_sqr, normalize,
                  renyi_entropy, von_neumann_entanglement_entropy, reduced_density_matrix, steiner_path, symmetric_gauge, symmetrize_and_normalize)
from ._lib import TnqsError, LIB_PATH  # noqa: F401
from .distributed import partition_vertices, cut_edges, shard  # noqa: F401
