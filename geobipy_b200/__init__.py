"""geobipy_b200 - B200-native per-sounding rjMCMC EM inversion path (drop-in for GeoBIPy's
Inference1D / FdemDataPoint / Model hot path).  See DESIGN.md."""
__version__ = "0.1.0"
