"""The namespace the reference looks classes up in (`getattr(arch, config[name]["name"])`,
libfewshot_core/utils/utils.py:20-35 with `arch = libfewshot_core.model`, trainer.py:15,438,454)."""
from ..backbone import BdcPool, Conv64F, resnet12, resnet12Bdc
from .abstract_model import AbstractModel, FinetuningModel, MetricModel, ModelType
from .deepbdc import DeepBDC
from .dn4 import DN4
from .maml import MAML, MetaModel, convert_maml_module
from .meta_baseline import MetaBaseline
from .proto_net import ProtoNet


def get_instance(module, name, config, **kwargs):
    """utils.get_instance (utils.py:20-35)."""
    if config[name]["kwargs"] is not None:
        kwargs.update(config[name]["kwargs"])
    return getattr(module, config[name]["name"])(**kwargs)


__all__ = ["AbstractModel", "MetricModel", "FinetuningModel", "ModelType", "MetaModel", "ProtoNet", "DN4", "DeepBDC", "MAML", "MetaBaseline", "convert_maml_module", "Conv64F", "resnet12",
           "resnet12Bdc", "BdcPool", "get_instance"]
