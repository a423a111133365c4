from pantea_b200.utils.tokenize import tokenize

__all__ = ["tokenize"]
