"""B200-native episodic few-shot audio hot path (drop-in for LibFewShot-audio's
ProtoNet / DN4 / DeepBDC set_forward / set_forward_loss path).  See DESIGN.md."""
__version__ = "0.1.0"
