"""empty stub: plotting is out of scope"""
