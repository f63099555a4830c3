"""Empty stub: boto3 is imported (never used) by layers/bert/file_utils.py:19."""
