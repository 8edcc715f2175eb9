"""aivc_b200 -- B200-native implementation of AIVC's per-frame encode/decode hot path."""
__version__ = '0.1.0'
