"""Logging like bhmm/util/logger.py:25-50: a 'BHMM' logger on stdout whose level follows config.verbose."""
import logging
import sys

from . import config

_logger = None


def logger(name='BHMM', pattern='%(asctime)s %(levelname)s %(name)s: %(message)s', date_format='%H:%M:%S',
           handler=None):
    global _logger
    if _logger is not None:
        return _logger
    _logger = logging.getLogger(name + '_b200')
    _logger.setLevel(logging.INFO if config.verbose else logging.WARNING)
    if not _logger.handlers:
        h = handler or logging.StreamHandler(sys.stdout)
        h.setFormatter(logging.Formatter(pattern, date_format))
        _logger.addHandler(h)
    return _logger
