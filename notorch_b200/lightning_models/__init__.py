from .glue import TensorDictModuleLite, TensorDictSequentialLite, prepare_graph, transfer_batch_to_device

__all__ = ["transfer_batch_to_device", "prepare_graph", "TensorDictModuleLite", "TensorDictSequentialLite"]
