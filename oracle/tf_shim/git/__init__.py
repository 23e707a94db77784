class Repo:  # utils/utils.py:108-111 only needs the import to succeed
    pass
