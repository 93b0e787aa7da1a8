class _Config:
    @staticmethod
    def set_visible_devices(devices, kind):
        return None


config = _Config()
