class SchedulerMixin:
    config_name = "scheduler_config.json"
