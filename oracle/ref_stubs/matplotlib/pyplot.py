# empty import stub (see README.md)
