"""Network-topology helpers with the reference's names and semantics (utils.py:87-130).
Host-side, constructor-time only."""
from __future__ import annotations

import numpy as np


def check_connections(connections):
    """utils.py:87-92 — a child index must not be smaller than its parent's."""
    for node, children in connections.items():
        for child in children or []:
            if child < node:
                raise Exception("Downstream node cannot have a smaller index number than upstream node")


def create_network(connections):
    """utils.py:94-102 — adjacency matrix, network[parent][child] = 1."""
    n = max(connections.keys()) + 1
    network = np.zeros((n, n))
    for parent, children in connections.items():
        for child in children or []:
            network[parent][child] = 1
    return network


def get_stage(node, network):
    """utils.py:105-120 — number of edges between ``node`` and node 0."""
    stage = 0
    hops = 0
    while node != 0:
        parents = np.nonzero(network[:, node])[0]
        if len(parents) == 0 or hops > len(network):
            raise Exception("Infinite Loop")
        node = int(parents[0])
        stage += 1
        hops += 1
    return stage


def get_retailers(network):
    """utils.py:124-130 — nodes without downstream nodes."""
    return [i for i in range(len(network)) if not np.any(network[i])]
