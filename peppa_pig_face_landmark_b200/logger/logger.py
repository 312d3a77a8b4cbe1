"""Package logger (the reference configures the root logger at import,
Skps/logger/logger.py:11-25; here a named logger is used and nothing global is touched)."""
import logging

logger = logging.getLogger("Skps")
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("[%(asctime)s] [%(levelname)s] %(message)s "))
    logger.addHandler(_h)
    logger.setLevel(logging.WARNING)
    logger.propagate = False
