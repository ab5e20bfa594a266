import logging as _logging


def get_logger(name=None):
    return _logging.getLogger(name)
