"""import-only stub (oracle scaffolding)"""
