class BaseData:
    pass
