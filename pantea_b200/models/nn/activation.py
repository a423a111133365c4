"""Activation names and their integer codes in the C ABI (table of reference
`pantea/models/nn/activation.py:48-60`; `exp` is exp(-x) as there)."""
ACTIVATION_CODES = {
    "identity": 0, "tanh": 1, "logistic": 2, "softplus": 3, "relu": 4,
    "gaussian": 5, "cos": 6, "exp": 7, "harmonic": 8,
}
