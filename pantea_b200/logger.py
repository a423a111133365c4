"""Logging convention of the reference: `logger.error(msg, exception=X)` logs then raises
(reference `pantea/logger.py:80-83`)."""
from __future__ import annotations

import logging
from typing import Optional, Type


class Logger:
    def __init__(self, name: str = "pantea_b200", level: int = logging.WARNING) -> None:
        self._log = logging.getLogger(name)
        if not self._log.handlers:
            handler = logging.StreamHandler()
            handler.setFormatter(logging.Formatter("%(levelname)s: %(message)s"))
            self._log.addHandler(handler)
        self._log.setLevel(level)

    def debug(self, msg: str) -> None:
        self._log.debug(msg)

    def info(self, msg: str) -> None:
        self._log.info(msg)

    def print(self, msg: str = "", **kwargs) -> None:
        print(msg, **kwargs)

    def warning(self, msg: str) -> None:
        self._log.warning(msg)

    def error(self, msg: str, exception: Optional[Type[BaseException]] = None) -> None:
        self._log.error(msg)
        if exception is not None:
            raise exception(msg)

    def set_level(self, level: int) -> None:
        self._log.setLevel(level)

    @property
    def level(self) -> int:
        return self._log.level


logger = Logger()


def set_logging_level(level: int) -> None:
    logger.set_level(level)


class LoggingContextManager:
    """Temporarily change the logging level (reference `pantea/logger.py:96-118`)."""

    def __init__(self, level: int) -> None:
        self.level = level
        self._old = logger.level

    def __enter__(self) -> "LoggingContextManager":
        self._old = logger.level
        logger.set_level(self.level)
        return self

    def __exit__(self, *exc) -> None:
        logger.set_level(self._old)
