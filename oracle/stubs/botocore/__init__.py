"""Empty stub: botocore is imported (never used) by layers/bert/file_utils.py:21."""
