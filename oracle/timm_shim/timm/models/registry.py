def get_pretrained_cfg(name):
    return {"name": name}
