"""stub"""
Image = None
