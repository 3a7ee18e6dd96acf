"""hdgnn_b200 -- B200-native hot path of HD-GNN (forward/backward of the graph2graph network)."""
__version__ = "0.1.0"
