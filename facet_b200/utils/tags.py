"""Tag string helpers and tagging parameters (utils/tags.py:8-52 of the reference)."""


def tags_to_string(tags):
    if not tags:
        return None
    return ",".join(tags)


def string_to_tags(tags_str):
    if not tags_str:
        return []
    return [t.strip() for t in tags_str.split(",") if t.strip()]


def get_tag_params(config):
    clip_settings = config.get_clip_settings()
    tag_settings = config.get_tagging_settings()
    return clip_settings.get("similarity_threshold_percent", 22) / 100, tag_settings.get("max_tags", 5)
