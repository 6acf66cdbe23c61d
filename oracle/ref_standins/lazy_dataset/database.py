class Database:
    pass


class JsonDatabase(Database):
    pass
