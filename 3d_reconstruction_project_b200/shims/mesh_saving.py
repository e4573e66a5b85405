"""Stand-in for the reference's mesh_saving.py (OUT OF SCOPE)."""


class MeshSaving:
    def save_mesh(self, mesh, densities, filename="output_mesh_on_the_fly.ply", colored_filename="colored_output_mesh_on_the_fly.ply"):
        print("mesh saving is outside the b200recon hot path: skipped")
