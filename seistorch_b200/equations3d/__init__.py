"""3D equation plug-ins."""
