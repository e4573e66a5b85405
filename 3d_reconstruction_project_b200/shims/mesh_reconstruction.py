"""Stand-in for the reference's mesh_reconstruction.py (Poisson meshing: OUT OF SCOPE)."""


class MeshReconstruction:
    def reconstruct_mesh(self, pcd, depth=6):
        print("mesh reconstruction is outside the b200recon hot path: skipped")
        return None, None
