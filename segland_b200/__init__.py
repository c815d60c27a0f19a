"""segland_b200 -- B200-native POP head + dense post-processing for SegLand."""
__version__ = "0.1.0"
