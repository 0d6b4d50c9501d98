from .Mesh import Mesh
