"""Line tokenizer used by the RuNNer-format readers (behaviour of reference `pantea/utils/tokenize.py:4-22`)."""
from __future__ import annotations

from typing import List, Optional, Tuple


def tokenize(line: str, comment: Optional[str] = None) -> Tuple[Optional[str], List[str]]:
    """Split `line` into a lower-cased keyword and the remaining tokens.

    With `comment` given, a line starting with it yields `(None, [])`; otherwise the line is cut
    at `line.find(comment)` -- which is -1 when no comment is present, i.e. the last character is
    dropped (the trailing newline when reading from a file), as the reference does.
    """
    if comment is not None:
        if line.startswith(comment):
            return None, []
        line = line[: line.find(comment)]
    parts = line.split()
    if not parts:
        return None, []
    return parts[0].lower(), parts[1:]
