mplot3d = None
